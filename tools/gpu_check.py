"""GPU bring-up helper: run the golden parity cases and print every residual (tools; not part of the product)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from parity import run_case_cuda_vs_oracle
from helpers import GOLDEN_CASES

names = sys.argv[1:] or GOLDEN_CASES
bad = 0
for n in names:
    try:
        res = run_case_cuda_vs_oracle(n, 'cuda:0')
    except Exception as e:   # keep going: the other cases narrow the failure down
        print(n, 'EXCEPTION', repr(e))
        bad += 1
        try:
            torch.cuda.synchronize()
        except Exception as e2:
            print('device is gone:', repr(e2))
            break
        continue
    print(n, 'OK' if res['ok'] else 'FAIL', {k: (f'{v:.2e}' if isinstance(v, float) else v) for k, v in res.items()})
    bad += 0 if res['ok'] else 1
print('failures:', bad)
sys.exit(1 if bad else 0)
