"""profiles/r02_bench_results.md from the bench lines kept under profiles/r02/ (copies of what gpurun calls left in gpurun_out/)."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "profiles", "r02")


def line(p):
    with open(p) as f:
        txt = [l for l in f.read().strip().splitlines() if l.startswith("{")]
    return json.loads(txt[-1])


rows = []
for p in sorted(glob.glob(os.path.join(D, "bench_*.json"))):
    name = os.path.basename(p)[6:-5]
    if name == "ref" or name.endswith("before_memo"):
        continue
    rows.append((name, line(p)))
order = {"c5": 0, "c4": 1, "c3": 2, "c2": 3}
rows.sort(key=lambda r: (order.get(r[0].split("_")[0], 9), r[1]["n_gpus"], r[0]))
out = ["# Round 2 — bench.py lines measured on B200 through gpurun (the JSON lines themselves: profiles/r02/bench_*.json)",
       "",
       "`value` = cells / device time per step (one device-side stopwatch over all K steps: kernels, epoch resets, the one host",
       "sync per step, idle gaps; max over ranks); `e2e` = host arrays -> C ABI -> decision on the host, wall clock.",
       "",
       "| file | workload | GPUs | cells | value (cells/s) | ms/step | e2e (cells/s) | dominant kernel | ms | frac of measured HBM peak | launches/step | SM MHz |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for name, r in rows:
    rf = r["roofline"]
    clk = r.get("clocks") or {}
    out.append("| %s | %s | %d | %d | %.3e | %.3f | %.3e | %s | %.3f | %.3f | %d | %s %s |" % (
        name, r["config"]["workload"].split(",")[0], r["n_gpus"], r["config"]["cells"], r["value"], r["ms_per_step"],
        r["e2e"]["value"], rf["kernel"].split(" [")[0], rf["kernel_ms"], rf["frac"],
        r["gpu_launches"] // max(1, r["steps"]), clk.get("sm_mhz"), clk.get("reasons")))
    rc = r.get("roofline_carried")
    if rc:
        out.append("| | carried path (the same batch as two halves, no clear(): every row read-modify-written) | %d | | | | | %s | %.3f | %.3f | | |" % (
            r["n_gpus"], rc["kernel"].split(" [")[0], rc["kernel_ms"], rc["frac"]))
ref = os.path.join(D, "bench_ref.json")
if os.path.exists(ref):
    r = line(ref)
    cb = r["cpu_baseline"]
    out += ["", "CPU arm (`bench.py --impl reference`, the oracle's literal restatement, %d threads): **%.3e cells/s** on its fixed sample "
            "(apply %.2f s + tally %.2f s); the whole-cluster extrapolation is a side field, never the value." % (
                cb["cores"], r["value"], cb["apply_s"], cb["tally_s"])]
out += ["", "Notes.  The c5_1gpu line is the shipped build (memo + L2 prefetch on the read-modify-write path, `profiles/r02_ab_carried.md`), run",
        "without the CPU leg; `bench_c5_1gpu_before_memo.json` is the full line (with `cpu_baseline`) of the build before it.  The 4-GPU lines predate",
        "that change (the fresh path they time is unchanged; their carried figure is the old 0.51).  2- and 8-GPU lines were measured earlier in the",
        "round (quoted in DESIGN.md §5: 5.17e7 cells/s at 1.93 ms, 1.30e8 at 0.76 ms) but their JSON files did not survive a container",
        "replacement; the driver's own SCALE run at round end is authoritative for 1 / 2 / 4 / 8 GPUs."]
with open(os.path.join(ROOT, "profiles", "r02_bench_results.md"), "w") as f:
    f.write("\n".join(out) + "\n")
print("\n".join(out))
