"""MPAS-layout adapter bandwidth against the number of tracers of the group (run length / alignment of the
MPAS side): is the SoA -> MPAS direction held back by runs that are not whole 128-byte lines?"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import torch  # noqa: E402

pkg = bench.ge.load_package(); host = pkg.host
nL, nC = 60, 235160
ctx = host.Context(nL, nC, device=0, parms=host.Parms())
st = torch.cuda.Stream(); ctx.set_stream(st.cuda_stream)
for nT in (8, 14, 16, 30, 32, 48):
    mpas = torch.rand((nC, nL, nT), dtype=torch.float64, device="cuda")
    soa = torch.empty((nT, nL, nC), dtype=torch.float64, device="cuda")
    slot = list(range(1, nT + 1))
    torch.cuda.synchronize()
    res = []
    for name, fn, passes in (("to_soa", lambda: ctx.mpas_to_soa(mpas.data_ptr(), soa.data_ptr(), slot, nL, nC), 2),
                             ("update", lambda: ctx.soa_to_mpas(soa.data_ptr(), mpas.data_ptr(), slot, nL, nC, alpha=1e-9, beta=1.0), 3),
                             ("convert", lambda: ctx.soa_to_mpas(soa.data_ptr(), mpas.data_ptr(), slot, nL, nC, alpha=1.0, beta=0.0), 2)):
        fn(); ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            fn()
        e1.record(st); ctx.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res.append("%s %.3f ms %.0f GB/s" % (name, ms, passes * nT * nL * nC * 8 / ms / 1e6))
    print("nT %2d (%4d B per level): %s" % (nT, nT * 8, "; ".join(res)), flush=True)
    del mpas, soa
