import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kernel_checks as K
from leastereo_b200.kernels import get_ops
ops = get_ops()
dev = torch.device("cuda:0")
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for i, case in enumerate(K.TC_CASES + K.TC_ROLL_CASES):
    try:
        w = K.check_conv_tc(ops, dev, planes=planes, cases=[case], verbose=False)
        torch.cuda.synchronize()
        print("case", i, case, "ok", w, "status", ops.lib.lea_tc_status(), flush=True)
    except AssertionError as e:
        print("case", i, case, "MISMATCH", e, "status", ops.lib.lea_tc_status(), flush=True)
    except Exception as e:
        print("case", i, case, "FAIL", str(e)[:100], "status", ops.lib.lea_tc_status(), flush=True)
        break
