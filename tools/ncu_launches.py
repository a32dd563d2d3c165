"""Turn the CSV of `ncu --metrics gpu__time_duration.sum --csv` into the launch list kept under profiles/.

    python tools/ncu_launches.py gpurun_out/launches.csv "command line that was profiled" > profiles/rNN_ncu_launches_bench.txt
"""
import csv
import io
import sys

UNIT_NS = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}


def main(path, command):
    rows = [line for line in open(path) if not line.startswith("==")]
    recs = list(csv.DictReader(io.StringIO("".join(rows))))
    ours = [x for x in recs if "ml::" in x["Kernel Name"] or "k_steric" in x["Kernel Name"]]
    out = [command,
           "(cold-cache, serialised launch times: compare SHARES, not absolutes). Kernels of libmomlevel_b200 only;",
           f"the other {len(recs) - len(ours)} launches in the capture are torch kernels generating the synthetic dataset "
           "before the timed region.", "",
           f"{'id':5s} {'kernel':78s} {'block':14s} {'grid':14s} gpu__time_duration.sum [ns]"]
    total = {}
    for x in ours:
        name = x["Kernel Name"].split("(CUtensorMap")[0].replace("ml::tma", "tma")
        ns = float(x["Metric Value"].replace(",", "")) * UNIT_NS.get(x["Metric Unit"], 1.0)
        out.append(f"{x['ID']:5s} {name[:78]:78s} {x['Block Size']:14s} {x['Grid Size']:14s} {ns:.0f}")
        total[name] = total.get(name, 0.0) + ns
    out += ["", "share of the captured launches of this library:"]
    s = sum(total.values()) or 1.0
    for k, v in total.items():
        out.append(f"  {k[:78]:78s} {100 * v / s:6.2f} %")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
