"""Per-launch counts of the dominant kernels from `ncu --set full` captures -> profiles/r2_kernel_counts.json, stamped with the digest
of the CUDA sources they were taken from (bench.py refuses the numbers when the sources have changed since).

  python tools/make_kernel_counts.py gpurun_out/prof_a.ncu-rep [gpurun_out/prof_b.ncu-rep ...]
Every captured launch of update_half_kernel / simulate_accept_kernel becomes one entry `<kernel>@<updates per launch>`; the updates
per launch are read from the kernel's own accept/launch geometry via the `--updates` list given in the same order as the reports."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_digest  # noqa: E402

NAMES = {"GaussMean": "gauss_mean", "SirTauLeap": "sir_tauleap", "Logistic": "logistic"}


def main():
    args = sys.argv[1:]
    out = {"csrc_sha256_16": csrc_digest(), "source": "ncu --set full --clock-control none, one launch per kernel; tools/make_kernel_counts.py", "kernels": {}}
    for spec in args:
        path, updates = spec.split(":")
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, data = rows[0], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        for r in data:
            name = r[col["Kernel Name"]]
            m = re.search(r"(update_half_kernel|simulate_accept_kernel)<(?:sabc::)?(\w+)", name)
            if not m:
                continue
            kind, model = m.group(1), NAMES.get(m.group(2), m.group(2))
            key = (f"{kind}<{model}, DE>" if kind == "update_half_kernel" else f"{kind}<{model}>") + f"@{updates}"
            f = lambda k: float(r[col[k]].replace(",", ""))                       # noqa: E731
            unit = lambda k: rows[1][col[k]]                                      # noqa: E731
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            dram = f("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")] + f("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
            out["kernels"][key] = {
                "dram_bytes_per_launch": dram, "warp_inst_per_launch": f("smsp__inst_executed.sum"),
                "thread_inst_per_launch": f("thread_inst_executed"),
                "issue_active_pct_under_ncu": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "l1_data_pipe_wavefronts_pct_under_ncu": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                "duration_us_under_ncu": f("gpu__time_duration.sum"), "registers": f("launch__registers_per_thread"),
                "report": os.path.basename(path)}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_kernel_counts.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
