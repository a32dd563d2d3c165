"""Wall time of the public call momlevel_b200.steric(dset) on a device-resident OM4p25 year, against the kernel alone."""
import cProfile, pathlib, pstats, sys, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import momlevel_b200 as ml
from momlevel_b200 import core, synth

nt, nz, ny, nx = synth.CONFIGS["om4p25"]
ds = synth.make_dataset(nt, nz, ny, nx, seed=123, device="cuda", dtype=torch.float32)
pts = nt * nz * ny * nx
def wall(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best
for variant in ("steric", "thermosteric", "halosteric"):
    t = wall(lambda: ml.steric(ds, variant=variant))
    print(f"ml.steric(ds, variant={variant!r}): {t * 1e3:.3f} ms wall  -> {pts / t / 1e9:.1f} G points/s")
res, ref = ml.steric(ds)
t = wall(lambda: ml.steric(ds, reference=ref))
print(f"ml.steric(ds, reference=ref): {t * 1e3:.3f} ms wall  -> {pts / t / 1e9:.1f} G points/s")
t = wall(lambda: ml.steric(ds, domain='global', reference=ref))
print(f"ml.steric(ds, domain='global', reference=ref): {t * 1e3:.3f} ms wall  -> {pts / t / 1e9:.1f} G points/s")
pres = ds["z_l"].data * 1e4 + 101325.0
t = wall(lambda: core.steric_local_selfref(ds["thetao"].data, ds["so"].data, ds["volcello"].data[0], ds["z_i"].data, ds["deptho"].data, pres))
print(f"core.steric_local_selfref: {t * 1e3:.3f} ms wall  -> {pts / t / 1e9:.1f} G points/s")
pr = cProfile.Profile(); pr.enable()
for _ in range(5): ml.steric(ds)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
