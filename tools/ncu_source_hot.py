"""Aggregate the ncu source page by CUDA-C source line: share of warp instructions, stall samples, average active threads.
usage: ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:<k> --launch-count 1 > src.csv
       python tools/ncu_source_hot.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr) or not r[0].strip().isdigit(): continue
    i_ie, i_te, i_s = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    try:
        out.append((float(r[i_ie]), float(r[i_te]), float(r[i_s]), cur, r[0], r[1]))
    except ValueError:
        pass
tot_i = sum(o[0] for o in out); tot_t = sum(o[1] for o in out); tot_s = sum(o[2] for o in out)
print(f"total warp-instr {tot_i:.4g}, thread-instr {tot_t:.4g}, avg active threads {tot_t / max(tot_i, 1):.2f}, samples {tot_s:.0f}")
print(f"{'file:line':24s} {'%instr':>7s} {'%samples':>8s} {'thr/instr':>9s}  source")
for ie, te, ss, f, ln, src in sorted(out, key=lambda o: -o[0])[:top]:
    print(f"{f + ':' + ln:24s} {ie / tot_i:7.2%} {ss / max(tot_s, 1):8.2%} {te / max(ie, 1):9.2f}  {src.strip()[:105]}")
