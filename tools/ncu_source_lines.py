"""Per-source-line instruction counts and stall samples from an .ncu-rep captured with --import-source on.
    python tools/ncu_source_lines.py gpurun_out/<tag>/blend.ncu-rep blend_bwd [top_n]
"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = ""
lines = []
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0] != "":
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[7] or 0), int(r[4] or 0)))
        except ValueError:
            pass
tot_i = sum(l[3] for l in lines); tot_s = sum(l[4] for l in lines)
print(f"# {rep} kernel~{pat}: {tot_i/1e6:.1f} M warp instructions, {tot_s} stall samples")
print(f"{'inst%':>6} {'smpl%':>6} {'Minst':>8}  file:line  source")
for f, n, src, i, s in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100*i/tot_i:6.2f} {100*s/max(tot_s,1):6.2f} {i/1e6:8.1f}  {f}:{n}  {src[:100]}")
