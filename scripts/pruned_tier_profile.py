import os, sys, torch
os.environ.setdefault("VQB200_EXPERIMENTAL", "1")
sys.path.insert(0, "/root/repo")
from vq_gan_b200 import ops
g = torch.Generator().manual_seed(1)
n = 1 << 20
centres = torch.randn(16, 256, generator=g)
E = (centres[torch.randint(0, 16, (16384,), generator=g)] + 1e-4 * torch.randn(16384, 256, generator=g)).cuda()
z = (centres[torch.randint(0, 16, (n,), generator=g)] + 0.05 * torch.randn(n, 256, generator=g))
z = z.view(1024, 1024, 256).permute(0, 2, 1).contiguous().view(1024, 256, 1024, 1).cuda()
for _ in range(2):
    idx, dmin, st = ops.search(z, E, 4)
torch.cuda.synchronize()
print(st.tolist())
