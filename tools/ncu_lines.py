"""Per-source-line summary of an ncu report: python tools/ncu_lines.py rep kernel_regex [top]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data = []
cur_file = ""
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    if len(r) >= 8 and r[0].isdigit() and r[2] == "-":
        try:
            data.append((cur_file, int(r[0]), r[1].strip(), int(r[6]), int(r[7])))
        except ValueError:
            pass
ts = sum(d[3] for d in data) or 1
ti = sum(d[4] for d in data) or 1
print(f"total samples {ts}, warp instructions {ti}")
for d in sorted(data, key=lambda x: -x[3])[:top]:
    print(f"{d[3] / ts * 100:5.1f}% smp {d[4] / ti * 100:5.1f}% inst  {d[0]}:{d[1]:<4d} {d[2][:100]}")
