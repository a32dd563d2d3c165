"""Where config 3's time goes: host time per member call, device time, and a CUDA-graph replay of the member loop."""
import pathlib, sys, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from momlevel_b200 import core, synth
from momlevel_b200 import distributed as mld

nt, nz, ny, nx = synth.CONFIGS["spear1deg"]
dev = torch.device("cuda")
grid = synth.make_grid(nz, ny, nx, seed=7, device=dev)
pres = (grid["z_l"] * 1.0e4 + 101325.0).contiguous()
z_i, depth = grid["z_i"].contiguous(), grid["deptho"].contiguous()
fields = [synth.make_fields(grid, nt, seed=1000 + m, dtype=torch.float32) for m in range(4)]
outs = [core.selfref_outputs(T, S) for T, S, _ in fields]
ev = lambda: torch.cuda.Event(enable_timing=True)

def run_all():
    for (T, S, V), o in zip(fields, outs):
        core.steric_local_selfref(T, S, V, z_i, depth, pres, out=o)

run_all(); torch.cuda.synchronize()
# host time of one call (GPU idle at the start), device time of the same call
for _ in range(2):
    torch.cuda.synchronize()
    a, b = ev(), ev()
    t0 = time.perf_counter(); a.record()
    core.steric_local_selfref(*fields[0], z_i, depth, pres, out=outs[0])
    b.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"one member: host {1e3 * (t1 - t0):.3f} ms, device {a.elapsed_time(b):.3f} ms")
for _ in range(3):
    torch.cuda.synchronize()
    a, b = ev(), ev()
    t0 = time.perf_counter(); a.record(); run_all(); b.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"4 members eager: host {1e3 * (t1 - t0):.3f} ms, device {a.elapsed_time(b):.3f} ms")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run_all()
g.replay(); torch.cuda.synchronize()
for _ in range(3):
    a, b = ev(), ev()
    a.record(); g.replay(); b.record()
    torch.cuda.synchronize()
    print(f"4 members graph replay: device {a.elapsed_time(b):.3f} ms")
want = [tuple(t.clone() for t in o) for o in outs]
for o in outs:
    for t in o: t.zero_()
g.replay(); torch.cuda.synchronize()
ok = all(torch.equal(torch.nan_to_num(x), torch.nan_to_num(y)) for o, w in zip(outs, want) for x, y in zip(o, w))
print("graph replay reproduces eager results:", ok)
