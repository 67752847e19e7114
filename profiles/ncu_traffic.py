"""DRAM traffic per sample of the rated kernels from an `ncu --set full` report -> JSON read by bench.py.

    python profiles/ncu_traffic.py gpurun_out/r02_full.ncu-rep SAMPLES_PER_LAUNCH RAYS_PER_LAUNCH > profiles/r02_ncu_traffic.json

`traffic` = dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (the last captured launch of each
kernel, i.e. the warm one), divided by the samples that launch processed."""
import csv, io, json, subprocess, sys

ENTRY = {"hashgrid_fwd_kernel": "den_hashgrid_fwd", "hashgrid_bwd_kernel": "den_hashgrid_bwd",
         "mlp_fwd_tc_kernel": "den_mlp_fwd", "mlp_bwd_tc_kernel": "den_mlp_bwd",
         "composite_fwd_kernel": "den_composite_fwd", "composite_bwd_sweep_kernel": "den_composite_bwd",
         "lpf_loss_kernel<0": "den_lpf_loss_fwd", "lpf_loss_kernel<1": "den_lpf_loss_bwd",
         "compact_kernel": "den_compact_samples_ex", "march_kernel<2>": "den_march_single"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}

rep, samples, rays = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
kernels = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    key = next((v for k, v in ENTRY.items() if k in name.replace("(bool)", "").replace("lpf::", "")), None)
    if key is None:
        continue
    def val(metric):
        return float(r[col[metric]]) * UNIT[units[col[metric]]]
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    kernels[key] = {
        "kernel": name.split("(")[0], "launch_id": int(r[0]), "ms_under_ncu": round(val("gpu__time_duration.sum"), 4),
        "dram_bytes_read": rd, "dram_bytes_write": wr,
        "dram_bytes_per_sample": (rd + wr) / samples,
        "registers": int(float(r[col["launch__registers_per_thread"]])),
        "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
        "tensor_pipe_active_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]),
        "l1tex_pct": float(r[col["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]]),
        "lts_pct": float(r[col["lts__throughput.avg.pct_of_peak_sustained_elapsed"]]),
        "dram_pct": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
        "warps_active_pct": float(r[col["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
    }
json.dump({"report": rep, "samples_per_launch": samples, "rays_per_launch": rays,
           "command": "ncu --set full --clock-control none --import-source on -k regex:... python bench.py "
                      "--steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-e2e (after the same command exited 0 "
                      "without ncu)", "kernels": kernels}, sys.stdout, indent=1)
