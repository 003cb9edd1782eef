import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import mpti_oracle as O
from r3dfsseg_b200 import ops
sd = torch.load('tests/golden/weights_fixture.pt')
g = torch.load('tests/golden/golden_dgcnn.pt')
x = g['dgcnn_x']; dev='cuda:0'
p='encoder.edge_convs.0'
def fold(pref):
    s = sd[pref+'.weight']/torch.sqrt(sd[pref+'.running_var']+1e-5)
    return s, sd[pref+'.bias']-sd[pref+'.running_mean']*s
s1,t1 = fold(p+'.layer.1'); s2,t2 = fold(p+'.layer.4')
W1 = sd[p+'.layer.0.weight'].reshape(64,18); W2 = sd[p+'.layer.3.weight'].reshape(64,64)
ref = O.edgeconv_block(x, sd, p, 20)
y, idx = ops.edgeconv(x.to(dev), W1.to(dev), s1.to(dev), t1.to(dev), W2.to(dev), s2.to(dev), t2.to(dev), 20, return_idx=True)
y = y.cpu(); idx = idx.cpu()
idx_ref = O.knn(x, 20)
print('idx set mismatch rows', int((idx.sort(-1)[0] != idx_ref.sort(-1)[0]).any(-1).sum()))
print('edgeconv err', float((y-ref).abs().max()/ref.abs().max()))
# linear K=9
xp = x.transpose(1,2).reshape(-1,9).contiguous()
wpq = torch.cat([W1[:,:9], W1[:,9:]-W1[:,:9]],0)
spq = torch.cat([s1,s1]); tpq = torch.cat([torch.zeros(64), t1])
PQ = ops.linear(xp.to(dev), wpq.to(dev), spq.to(dev), tpq.to(dev), 0).cpu()
PQ_ref = (xp @ wpq.t())*spq + tpq
print('linear K=9 err', float((PQ-PQ_ref).abs().max()/PQ_ref.abs().max()))
for K in (8, 16, 17, 64, 100):
    a = torch.randn(300, K); w = torch.randn(70, K)
    r = ops.linear(a.to(dev), w.to(dev), None, None, 0).cpu()
    print('linear K', K, float((r - a@w.t()).abs().max()))
# per-channel error
print('per-channel max err', ((y-ref).abs().amax(dim=(0,2)))[:16])
print('---- DGCNN module vs golden')
from r3dfsseg_b200.models import MPTI_SelfAtten
from r3dfsseg_b200.episodes import default_args, make_episode
m = MPTI_SelfAtten(default_args(2,1)); m.load_state_dict(sd); m = m.to(dev).eval()
l1, l2 = m.encoder(x.to(dev))
for name, got, rf in (('l1', l1.cpu(), g['dgcnn_l1']), ('l2', l2.cpu(), g['dgcnn_l2'])):
    e = (got-rf).abs().amax(1)/rf.abs().max()   # per point
    print(name, 'max', float(e.max()), 'frac pts > 1e-4', float((e>1e-4).float().mean()), 'median', float(e.median()))
outs,_ = (lambda enc: (enc.__setattr__('return_edgeconvs', True), enc(x.to(dev)))[1])(m.encoder)
m.encoder.return_edgeconvs = False
ref_outs = []
xx = x
for i in range(3):
    xx = O.edgeconv_block(xx, sd, f'encoder.edge_convs.{i}', 20); ref_outs.append(xx)
for i in range(3):
    e = (outs[i].cpu()-ref_outs[i]).abs().amax(1)/ref_outs[i].abs().max()
    print('edgeconv', i, 'max', float(e.max()), 'frac pts > 1e-4', float((e>1e-4).float().mean()))
    if i < 2:
        a = ops.knn(outs[i], 20).cpu(); b = O.knn(ref_outs[i], 20)
        print('   next-layer knn rows with different sets', int((a.sort(-1)[0] != b.sort(-1)[0]).any(-1).sum()))
print('---- features on a synthetic cloud')
ep = make_episode(5,2,1); xq = ep.query_x
rf = O.get_features(xq, sd); got = m.getFeatures(xq.to(dev)).cpu()
for lo,hi in ((0,64),(64,128),(128,192)):
    e = (got[:,lo:hi]-rf[:,lo:hi]).abs().amax(1)/rf[:,lo:hi].abs().max()
    print((lo,hi),'max', float(e.max()), 'frac pts > 2e-4', float((e>2e-4).float().mean()), 'median', float(e.median()))
a = ops.knn(xq.to(dev), 20).cpu(); b = O.knn(xq, 20)
print('layer-0 knn rows with different sets', int((a.sort(-1)[0] != b.sort(-1)[0]).any(-1).sum()), 'of', a.shape[0]*a.shape[1])
