"""What predict.py / evaluation.py of the reference look like on this framework, with
synthetic 8-bit images (no dataset in the sandbox): 8-bit pairs are uploaded double-buffered, z-normalised and padded
on the GPU (predict.py:144-184), run through the drop-in LEAStereo module, and scored on the device
(utils/metrics.py) against a synthetic ground truth.  Prints pairs/s from pinned 8-bit host images."""
import contextlib, io, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leastereo_b200 import LEAStereo, default_args  # noqa: E402
from leastereo_b200.pipeline import InputPipeline, disparity_metrics  # noqa: E402


def main(n_pairs=24, H=375, W=1242, crop_h=384, crop_w=1248, maxdisp=192):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = LEAStereo(default_args(maxdisp=maxdisp, cuda=True), dev).to(dev).eval()      # predict.py:52-55
    model.engine_options = {"assume_frozen": True}
    rng = np.random.RandomState(0)
    pairs = [(rng.randint(0, 256, (H, W, 3)).astype(np.uint8), rng.randint(0, 256, (H, W, 3)).astype(np.uint8))
             for _ in range(4)]
    target = (torch.rand(1, crop_h, crop_w) * maxdisp * 0.5).to(dev)
    pipe = InputPipeline(H, W, crop_h, crop_w, dev, depth=2)
    stats = []
    with torch.no_grad():
        pipe.submit(*pairs[0])
        for k in range(3):                                  # warm-up (plans, weight images)
            pipe.submit(*pairs[(k + 1) % 4]); l, r = pipe.next(); model(l, r)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(n_pairs):
            if k + 1 < n_pairs:
                pipe.submit(*pairs[(k + 1) % 4])            # next pair's 8-bit upload overlaps this pair's kernels
            left, right = pipe.next()                       # predict.py:144-184 on the device
            disp = model(left, right)                       # predict.py:192
            stats.append(disparity_metrics(disp, target, maxdisp))      # utils/metrics.py, one 56-byte read-back
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(json.dumps({"pairs": n_pairs, "pairs_per_s_from_8bit_host_images": round(n_pairs / dt, 2),
                      "h2d_bytes_per_pair": pipe.h2d_bytes_per_pair, "image": [H, W], "padded_to": [crop_h, crop_w],
                      "last_metrics": {k: (round(v, 4) if isinstance(v, float) else v) for k, v in stats[-1].items()}}))


if __name__ == "__main__":
    main()
