"""Summarise an .ncu-rep (read here, no GPU): per-kernel headline counters + top stall reasons + hottest SASS."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_src_fp64.sum.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"]
for r in rows[2:]:
    print("=" * 100)
    for i, h in enumerate(hdr):
        if h in want:
            print(f"{h:75s} {units[i]:12s} {r[i]}")
    st = []
    for i, h in enumerate(hdr):
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try: st.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError: pass
    st.sort(reverse=True)
    tot = sum(v for v, _ in st) or 1
    print("stalls:", ", ".join(f"{n} {100*v/tot:.1f}%" for v, n in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
ks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for a, b in zip(ks[:-1], ks[1:]):
    h = rows[a + 1]; body = rows[a + 2:b]
    isrc, isamp = h.index("Source"), h.index("# Samples")
    tot = sum(int(r[isamp]) for r in body) or 1
    print("=" * 100); print(rows[a][1][:110], "total samples", tot)
    hot = sorted(range(len(body)), key=lambda j: -int(body[j][isamp]))[:top]
    for j in sorted(hot):
        print(f"{j:6d} {100*int(body[j][isamp])/tot:6.2f}%  {body[j][isrc].strip()[:80]}")
