"""profiling target: examples/Arithmetic.hs `mul_` (n = 256, limbs 2 -> 2 -> 1) on 32768 ciphertext pairs"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, alchemy_b200 as ab
from alchemy_b200 import she
Q = [268440577, 8392193, 1073750017]
plan = ab.Plan(512, Q); n = plan.n; B = 32768
rnd = lambda lead, L: torch.stack([torch.randint(0, q, tuple(lead) + (n,), dtype=torch.int64, device="cuda") for q in Q[:L]], dim=-2)
a, b, hint = rnd((B, 2), 2), rnd((B, 2), 2), rnd((2, 2), 2)
for _ in range(2): she.mulKeySwitch(plan, a, b, hint, 2, 1, prescale=[pow(7, -1, q) for q in Q[:2]], dec_mask=1)
torch.cuda.synchronize()
