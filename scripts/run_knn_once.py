"""Tiny driver for ncu captures: a few knn launches at the episode shape (B clouds x 2048 pts)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r3dfsseg_b200 import ops
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 300
x = torch.randn(B, 2048, C, device="cuda").transpose(1, 2)
for _ in range(3):
    idx = ops.knn(x, 20)
torch.cuda.synchronize()
print(idx.shape)
