"""Tiny driver for ncu captures: two calls of E episodes through the C ABI (first = warm-up)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r3dfsseg_b200.episodes import default_args, make_episode
from r3dfsseg_b200.models import MPTI_SelfAtten
E = int(sys.argv[1]) if len(sys.argv) > 1 else 25
sd = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "weights_fixture.pt"))
m = MPTI_SelfAtten(default_args(2, 5)); m.load_state_dict(sd); m = m.cuda().eval()
eps = [make_episode(100 + i, 2, 5) for i in range(E)]
sx = torch.stack([e.support_x for e in eps]).cuda(); sy = torch.stack([e.support_y for e in eps]).cuda()
qx = torch.stack([e.query_x for e in eps]).cuda(); qy = torch.stack([e.query_y for e in eps]).cuda()
for _ in range(2):
    out = m.forward_episodes(sx, sy, qx, qy, eval=True)
torch.cuda.synchronize()
print(float(out["loss"].mean()))
