"""Three heights of an OM4p25 year: one-pass kernel (csrc/ml_tma3.cu) against one launch per height.

    python tools/variants_probe.py            (one B200; MOMLEVEL_B200_LIB selects an experiment build)

Prints one JSON line: ms per call by CUDA events (best of 5 after a warm-up) for the unfused path
(``ml_set_force_direct(2)``) and for the one-pass kernel at every chunk width, each with and without the
reference density stored, plus the largest difference between the two sets of heights.
"""

import json
import os
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from momlevel_b200 import core, synth  # noqa: E402


def best_ms(fn, n=5):
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    nt, nz, ny, nx = synth.CONFIGS["om4p25"]
    if len(sys.argv) > 1:
        nt = int(sys.argv[1])
    dev = torch.device("cuda", 0)
    grid = synth.make_grid(nz, ny, nx, seed=123, device=dev)
    T, S, V = synth.make_fields(grid, nt, seed=123, dtype=torch.float32)
    pres = (grid["z_l"] * 1.0e4 + 101325.0).contiguous()
    z_i, depth = grid["z_i"].contiguous(), grid["deptho"].contiguous()
    pts = nt * nz * ny * nx
    out = {"lib": os.environ.get("MOMLEVEL_B200_LIB", "default"), "nt": nt, "points": pts}

    def call(store):
        return core.steric_local_variants(T, S, V, z_i, depth, pres, want_rho_ref=store)

    core.force_direct(2)
    base = call(True)[0]
    out["three_launches_ms"] = best_ms(lambda: call(True))
    core.force_direct(0)
    for tile in (256, 128):
        for tc in (4, 6, 8, 12):
            core.variants_chunk(tc + (100 if tile == 128 else 200))
            got = call(False)[0]
            err = max(float(torch.nan_to_num(got[v] - base[v]).abs().max()) for v in got)
            ms = best_ms(lambda: call(False))
            out[f"one_pass_tile{tile}_tc{tc}"] = {"ms": ms, "ms_rho_ref_stored": best_ms(lambda: call(True)),
                                                  "max_abs_diff_vs_three_launches_m": err, "gpts": pts / ms / 1e6}
    core.variants_chunk(0)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
