"""Turn gpurun_out/launches.csv + *.ncu-rep into the committed text summaries under profiles/."""
import collections, csv, subprocess, sys

def launch_list(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[hdr_i], rows[hdr_i + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        name = r[ki].split("(")[0].replace("void ", "")[:64]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# command: python bench.py --steps 1 --warmup 1 --no-cpu --coalitions 152 ; launches {len(data)} ; total {tot:.0f} us\n")
        f.write(f"{'us':>12} {'share':>7} {'n':>5}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:12.1f} {100 * v[1] / tot:6.1f}% {v[0]:5d}  {k}\n")

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__cluster_dim_x", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]

def full(path, out, limit=64):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, from {path.split('/')[-1]} (one block per captured launch)\n")
        for r in rows[2:2 + limit]:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
            f.write(f"---- {d.get('Kernel Name', '')[:90]}\n")
            for w in WANT:
                if w in d and d[w] != "":
                    f.write(f"  {w} = {d[w]} {u[w]}\n")

if __name__ == "__main__":
    tag = sys.argv[1]
    launch_list("gpurun_out/launches.csv", f"profiles/{tag}_ncu_launch_list.txt")
    full("gpurun_out/prof_gemm_pair.ncu-rep", f"profiles/{tag}_ncu_full_gemm_pair.txt")
    full("gpurun_out/prof_others.ncu-rep", f"profiles/{tag}_ncu_full_other_kernels.txt")
