"""one call of each MPAS-layout adapter direction on the bench mesh (for ncu)"""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import torch  # noqa: E402
pkg = bench.ge.load_package(); host = pkg.host
nL, nC, nT = 60, 235160, 30
ctx = host.Context(nL, nC, device=0, parms=host.Parms())
mpas = torch.rand((nC, nL, nT), dtype=torch.float64, device="cuda")
soa = torch.empty((nT, nL, nC), dtype=torch.float64, device="cuda")
slot = list(range(1, nT + 1))
torch.cuda.synchronize()
for _ in range(2):
    ctx.mpas_to_soa(mpas.data_ptr(), soa.data_ptr(), slot, nL, nC)
    ctx.soa_to_mpas(soa.data_ptr(), mpas.data_ptr(), slot, nL, nC, alpha=1e-9, beta=1.0)
    ctx.soa_to_mpas(soa.data_ptr(), mpas.data_ptr(), slot, nL, nC, alpha=1.0, beta=0.0)
ctx.synchronize()
