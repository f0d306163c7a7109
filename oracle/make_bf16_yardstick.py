"""bf16 yard-stick: error of the reference math under torch.autocast(bfloat16) (CPU) against the fp64 golden
fixtures, per module case -> tests/golden/bf16_yardstick.json.  The reference has no bf16 mode of its own;
SURVEY.md §8d defines "bf16 parity" as no worse than 2x the reference under autocast.  Runs anywhere (uses the
oracle restatement, which tests pin bit-tight to the reference):  python oracle/make_bf16_yardstick.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import helpers as H  # noqa: E402
from oracle import gcn_oracle as O  # noqa: E402

only = [a for a in sys.argv[1:] if not a.startswith('-')]
path = os.path.join(ROOT, 'tests', 'golden', 'bf16_yardstick.json')
out = {}
if only and os.path.exists(path):
    with open(path) as f:
        out = json.load(f)
for name, case in H.CASES.items():
    if only and name not in only:
        continue
    built, fx = H.build_case(case), H.load_fixture(name)
    p = O.clone_state(built['state'], torch.float32, requires_grad=True)
    x = built['x'].clone().requires_grad_(True)
    extra = {k: v.clone().requires_grad_(True) for k, v in built['extra'].items()}
    with torch.autocast('cpu', dtype=torch.bfloat16):
        y = H.oracle_forward(case, x, p, extra)
    y.float().backward(built['cot'])
    pre = '' if case['kind'].endswith('_model') else 'm.'
    ks = [k for k in fx['grads'] if not k.startswith('__')]
    ks = [k for k in ks if p[pre + k].grad is not None]
    ga = torch.cat([p[pre + k].grad.reshape(-1) for k in ks])
    gb = torch.cat([fx['grads'][k].reshape(-1) for k in ks])
    gk = {k: O.rel_err(p[pre + k].grad, fx['grads'][k]) for k in ks if float(fx['grads'][k].norm()) > 0}
    out[name] = dict(y=O.rel_err(y, fx['y']), dx=O.rel_err(x.grad, fx['dx']), gall=O.rel_err(ga, gb), g=gk)
    print(name, {k: v for k, v in out[name].items() if k != 'g'}, 'worst g', max(gk.values()) if gk else 0)
with open(path, 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True)
