"""Summarises an `ncu --set full` report (run here, on the CPU box):
   python tools/ncu_full_summarize.py gpurun_out/r01_full.ncu-rep profiles/r01_ncu_full_summary
writes <out>.txt (one line per launch) and <out>.json (read by bench.py for `roofline.traffic`)."""
import csv
import io
import json
import re
import subprocess
import sys

HBM_PEAK_GBS = 6574.5      # MEASURED_PEAKS.json (copy bandwidth on this pool's B200s)
WANT = {
    "gpu__time_duration.sum": "dur",
    "sm__inst_executed_pipe_tensor.sum.pct_of_peak_sustained_active": "tensor_inst_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct_h",
    "dram__bytes_read.sum": "dram_r",
    "dram__bytes_write.sum": "dram_w",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "xbar_read",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "smem",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "smsp__inst_executed.sum": "inst",
}


def scale(v, unit):
    v = float(v.replace(",", "")) if isinstance(v, str) else float(v)
    u = unit.lower()
    for k, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0), ("msecond", 1e3), ("usecond", 1.0),
                 ("nsecond", 1e-3), ("second", 1e6)):
        if u.startswith(k):
            return v * m
    return v


def main(rep, out):
    if rep.endswith(".csv"):       # already exported with `ncu -i x.ncu-rep --page raw --csv`
        raw = "".join(l for l in open(rep) if l.startswith('"'))
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    res, lines = [], []
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        name = re.sub(r"void |<unnamed>::|\(anonymous namespace\)::", "", name)
        d = {"kernel": name, "grid": r[col["Grid Size"]]}
        for m, key in WANT.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                d[key] = scale(r[col[m]], units[col[m]])
        tensor = d.get("tensor_pct", d.get("tensor_pct_h", d.get("tensor_inst_pct", float("nan"))))
        dram = d.get("dram_r", 0.0) + d.get("dram_w", 0.0)
        l2sm = d.get("xbar_read", 0.0)     # bytes the SMs pulled over the crossbar (TMA loads included)
        res.append({"kernel": name, "grid": d["grid"], "us": d.get("dur", float("nan")), "tensor_pct": tensor,
                    "dram_bytes": dram, "l2_to_sm_bytes": l2sm, "l2_hit_pct": d.get("l2_hit"),
                    "regs": d.get("regs"), "smem_bytes": d.get("smem"), "warps_active_pct": d.get("occ_pct")})
        gbs = dram / (d.get("dur", float("nan")) * 1e-6) / 1e9 if d.get("dur") else float("nan")
        res[-1]["dram_gbs"] = gbs
        lines.append(f"{name[:44]:44s} grid {d['grid']:>16s} {d.get('dur', 0):8.1f} us  tensor-pipe {tensor:5.1f}%  "
                     f"dram {d.get('dram_r', 0) / 1e6:7.1f}+{d.get('dram_w', 0) / 1e6:6.1f} MB = {gbs:6.0f} GB/s ({100 * gbs / HBM_PEAK_GBS:4.1f}% of the measured {HBM_PEAK_GBS:.0f} GB/s)  "
                     f"L2->SM {l2sm / 1e6:7.1f} MB  L2 hit {d.get('l2_hit', 0):4.1f}%  regs {int(d.get('regs', 0))}  "
                     f"smem {d.get('smem', 0) / 1e3:.1f} KB  warps active {d.get('occ_pct', 0):4.1f}%")
    open(out + ".json", "w").write(json.dumps(res, indent=1))
    open(out + ".txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
