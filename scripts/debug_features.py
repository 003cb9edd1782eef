import os, sys, torch
sys.path.insert(0, '/root/repo')
from oracle import mpti_oracle as O
from r3dfsseg_b200.episodes import default_args, make_episode
from r3dfsseg_b200.models import MPTI_SelfAtten
sd = torch.load('/root/repo/tests/golden/weights_fixture.pt')
ep = make_episode(0, 2, 1)
m = MPTI_SelfAtten(default_args(2, 1)); m.load_state_dict(sd); m = m.cuda().eval()
x = torch.cat([ep.support_x.reshape(2, 9, -1), ep.query_x], 0)
with torch.no_grad():
    ref = O.get_features(x, sd)
got = m.getFeatures(x.cuda()).cpu()
rel = (got - ref).abs() / ref.abs().max()
print('feature rel err: max', float(rel.max()), 'median', float(rel.median()), 'frac > 1e-4:', float((rel.max(1)[0] > 1e-4).float().mean()), 'frac > 1e-3:', float((rel.max(1)[0] > 1e-3).float().mean()))
for name, sl in (('level1', slice(0, 64)), ('att', slice(64, 128)), ('base', slice(128, 192))):
    r = rel[:, sl]
    print(name, 'max', float(r.max()), 'pts >1e-4', float((r.max(1)[0] > 1e-4).float().mean()))
