"""profiles/r02_kernels.json from `ncu --set full` raw CSV exports: per kernel the DRAM bytes per launch, the tensor-pipe
activity and a few stall facts, with the export they came from.  bench.py reads the JSON for roofline.traffic.
   python scripts/ncu_to_profile.py gpurun_out/<tag>_full_raw.csv [more.csv ...]"""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"dx_red_kernel": "dx_moments", "dw_xf_kernel": "dw_top", "layer0_k": "layer0", "tc_gemm_kernel<(int)0, (int)0, (int)1": "fwd",
        "tc_gemm_kernel<0, 0, 1": "fwd", "out_backward": "out_backward(r01 path)", "image_col_reduce": "image_col_reduce(r01 path)"}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    out_path = os.path.join(ROOT, "profiles", "r02_kernels.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[col["Kernel Name"]]
            key = next((v for k, v in KEYS.items() if k in name), None)
            if key is None:
                continue
            def get(metric, scale_unit=None):
                if metric not in col:
                    return None
                v, u = num(r[col[metric]]), units[col[metric]]
                if v is None:
                    return None
                mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
                return v * mult
            rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
            out[key] = {
                "kernel": name[:120],
                "dram_bytes": (rd or 0) + (wr or 0), "dram_bytes_read": rd, "dram_bytes_write": wr,
                "duration_us_under_ncu": get("gpu__time_duration.sum"),
                "tensor_pipe_pct": get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "tc_pipe_pct": get("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active"),
                "xu_pipe_pct": get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                "issue_slots_pct": get("sm__inst_issued.avg.pct_of_peak_sustained_active"),
                "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
                "registers_per_thread": get("launch__registers_per_thread"),
                "source": "ncu --set full --clock-control none, " + os.path.basename(path),
            }
    json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, {kk: (round(vv, 1) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ("kernel", "source")})


if __name__ == "__main__":
    main()
