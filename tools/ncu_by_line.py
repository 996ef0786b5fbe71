"""Per-source-line view of an `ncu --set full --import-source on` capture without the GUI: joins the SASS page of the report
(`ncu -i REP --page source --csv`) with the line table of the same kernel in the built library (`nvdisasm -g`), instruction by
instruction, and sums stall samples, executed instructions and shared-memory wavefronts per line of engine.cu.
  python tools/ncu_by_line.py gpurun_out/r2_k_expand_full.ncu-rep [mangled-name fragment, default k_expandILi20ELi4ELi0E] [top N]
Runs where ncu / cuobjdump / nvdisasm are installed; no GPU needed.  The library must be the one that was profiled."""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "halo2-dynamic-sha256_b200", "libh2sha_b200.so")


def sass_lines(fragment):
    """[(opcode text, file line)] of the kernel whose mangled name contains `fragment`, in address order."""
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=d, check=True, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(d) if f.startswith("engine.") and f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=d, capture_output=True, text=True).stdout.splitlines()
    out, inside, line = [], False, 0
    for s in dis:
        if s.startswith(".text."):
            inside = fragment in s
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', s)
        if m:
            if m.group(1).endswith("engine.cu"):
                line = int(m.group(2))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?)\s*;", s)
        if m:
            out.append((m.group(2), line))
    return out


def main():
    rep = sys.argv[1]
    frag = sys.argv[2] if len(sys.argv) > 2 else "k_expandILi20ELi4ELi0E"
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    sass = sass_lines(frag)
    if len(sass) != len(data):
        print(f"warning: {len(sass)} instructions in the library vs {len(data)} in the report (different build?)", file=sys.stderr)
    keys = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive"]
    agg = {}
    for (op, line), r in zip(sass, data):
        a = agg.setdefault(line, [0.0] * len(keys))
        for k, name in enumerate(keys):
            try:
                a[k] += float(r[ix[name]])
            except ValueError:
                pass
    tot = [sum(a[k] for a in agg.values()) for k in range(len(keys))]
    src = open(os.path.join(ROOT, "halo2-dynamic-sha256_b200", "csrc", "engine.cu")).read().splitlines()
    print(f"# {rep}: totals samples {tot[0]:.0f}, warp instructions {tot[1]:.0f}, shared wavefronts {tot[2]:.0f} (excessive {tot[3]:.0f})")
    print(f"# {'line':>5} {'samples%':>8} {'inst%':>6} {'smem wf%':>8} {'excess':>9}  source")
    for line, a in sorted(agg.items(), key=lambda t: -t[1][0])[:top]:
        text = src[line - 1].strip()[:110] if 0 < line <= len(src) else ""
        print(f"  {line:5d} {100 * a[0] / tot[0]:8.2f} {100 * a[1] / tot[1]:6.2f} {100 * a[2] / max(tot[2], 1):8.2f} {a[3]:9.0f}  {text}")


if __name__ == "__main__":
    main()
