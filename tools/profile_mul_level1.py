import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, alchemy_b200 as ab
from alchemy_b200 import she
Q = [1543651201, 689270401, 718099201, 720720001, 1556755201, 1567238401]
plan = ab.Plan(20475, Q); n = plan.n; B = 512
rnd = lambda lead, L: torch.stack([torch.randint(0, q, tuple(lead) + (n,), dtype=torch.int64, device="cuda") for q in Q[:L]], dim=-2)
a, b, hint = rnd((B, 2), 4), rnd((B, 2), 4), rnd((5, 2), 5)
for _ in range(2): she.mulKeySwitch(plan, a, b, hint, 5, 3, prescale=[pow(32, -1, q) for q in Q[:4]], dec_mask=1)
torch.cuda.synchronize()
