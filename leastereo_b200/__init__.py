"""leastereo_b200 - B200-native hot path (cost volume -> 3D MatchingNet -> disparity head) behind the
reference's ``LEAStereo(args, device).forward(left, right)`` API.  See DESIGN.md."""
from .args import LEAStereoArgs, default_args, write_shipped_arch, SHIPPED_ARCH
from .leastereo import LEAStereo, Disp, DisparityRegression
from .modules import newFeature, newMatching, ConvBR3d
from .structure import network_layer_to_space

__all__ = ["LEAStereo", "LEAStereoArgs", "default_args", "write_shipped_arch", "SHIPPED_ARCH", "Disp",
           "DisparityRegression", "newFeature", "newMatching", "ConvBR3d", "network_layer_to_space"]
