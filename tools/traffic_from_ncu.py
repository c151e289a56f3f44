#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` raw csv (ncu -i X.ncu-rep --page raw --csv): per bench stage the DRAM
bytes (dram__bytes_read.sum + dram__bytes_write.sum), the L2 bytes (lts__t_sectors.sum x 32 B), the duration and the
tensor-pipe activity of its kernels, summed over the launches of ONE step.
usage: traffic_from_ncu.py raw.csv workload [out.json]"""
import csv, json, re, sys

STAGE = [  # kernel-name regex -> bench.py stage (api.cu StageTimer names)
    (r"k_density_select", "density_select"), (r"k_appearance_slab|k_appearance<0>|k_appearance<false>", "appearance_gather"),
    (r"k_scatter_walk<\d, ?(0|false)[,>]", "density_scatter"), (r"k_scatter_walk<\d, ?(1|true)[,>]", "appearance_scatter"),
    (r"k_fused_pack", "mlp_pack"), (r"k_mlp_fused_fwd", "mlp_fwd"), (r"k_mlp_fused_bwd|k_mlp_fused_wgrad|k_wgrad_reduce|k_fused_amax", "mlp_bwd"),
    (r"k_tc_rowgemm<3|k_encode_fwd|k_pack_all", "mlp_fwd"), (r"k_tc_rowgemm<2|k_tc_redgemm|k_encode_bwd|k_out_bwd", "mlp_bwd"),
    (r"k_composite_fwd", "composite"), (r"k_ray_bwd", "ray_bwd"), (r"k_pack<0", "pack"), (r"k_pack<1", "unpack"),
]


def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "sector": 1, "": 1, "%": 1}.get(u, 1)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return 0.0
        return float(r[i].replace(",", "")) * unit_scale(units[i])

    out = {}
    for r in body:
        name = r[col["Kernel Name"]]
        stage = next((s for pat, s in STAGE if re.search(pat, name)), None)
        if stage is None:
            continue
        e = out.setdefault(stage, {"dram_bytes": 0.0, "l2_bytes": 0.0, "ncu_us": 0.0, "launches": 0, "tensor_pct_time_weighted": 0.0})
        t = val(r, "gpu__time_duration.sum")
        e["dram_bytes"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        e["l2_bytes"] += 32.0 * val(r, "lts__t_sectors.sum")
        e["ncu_us"] += t
        e["launches"] += 1
        e["tensor_pct_time_weighted"] += t * val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    for e in out.values():
        e["tensor_pct_time_weighted"] = round(e["tensor_pct_time_weighted"] / max(e["ncu_us"], 1e-9), 2)
        e["ncu_us"] = round(e["ncu_us"], 2)
    path = sys.argv[3] if len(sys.argv) > 3 else "profiles/traffic.json"
    try:
        doc = json.load(open(path))
    except (OSError, ValueError):
        doc = {}
    doc[sys.argv[2]] = out
    doc["source"] = f"tools/traffic_from_ncu.py {sys.argv[1]} (ncu --set full, one step; per launch cold-cache and serialised)"
    json.dump(doc, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
