#!/usr/bin/env python
"""One eager step of MobileNetV1-ImageNet SFP-7, batch 256 (BASELINE.json configs[4], the depthwise path) between
cudaProfilerStart/Stop, for an ncu launch list with DRAM bytes per launch:

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/mbv1_launches.csv python tools/profile_mobilenet.py
    python tools/agg_launches.py gpurun_out/mbv1_launches.csv            # by kernel
    python tools/profile_mobilenet.py --table gpurun_out/mbv1_launches.csv   # launch by launch, GB/s against the HBM peak
"""
import csv, json, os, re, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def table(path):
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peak = 6536.7
    try:
        peak = float(json.load(open(peaks)).get("hbm_gbs", peak))
    except Exception:
        pass
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tunit = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
    rows = {}
    for row in csv.DictReader(l for l in open(path) if not l.startswith("==")):
        r = rows.setdefault(int(row["ID"]), {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("slfp::", "")[:70],
                                             "grid": row.get("Grid Size", ""), "us": 0.0, "bytes": 0.0})
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Name"] == "gpu__time_duration.sum":
            r["us"] += v * tunit.get(row["Metric Unit"].strip(), 1.0)
        else:
            r["bytes"] += v * unit.get(row["Metric Unit"].strip(), 1)
    print(f"| # | kernel | grid | us | DRAM MB | GB/s | of HBM peak ({peak:.0f}) |\n|---|---|---|---|---|---|---|")
    for i, r in sorted(rows.items()):
        gbs = r["bytes"] / r["us"] / 1e3 if r["us"] else 0.0
        print(f"| {i} | `{r['name']}` | {r['grid']} | {r['us']:.1f} | {r['bytes'] / 1e6:.1f} | {gbs:.0f} | {gbs / peak:.2f} |")
    print(f"| | **total** | | {sum(r['us'] for r in rows.values()):.1f} | {sum(r['bytes'] for r in rows.values()) / 1e6:.0f} | | |")


def main():
    import numpy as np, torch
    from cnns_slfp_quantization_b200 import engine, nets_common as nc, calibration
    from cnns_slfp_quantization_b200.nets_imgnet import MobileNetV1_Q
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    batch, size = 256, 224
    m32 = MobileNetV1_Q(3, 32).eval()
    sd = nc.synth_state_dict(m32)
    m32.load_state_dict(sd)
    n_layers = len(nc.quantized_layers(m32))
    nc.set_scales(m32, np.ones(n_layers), np.ones(n_layers))
    m32 = m32.to(dev)
    ka, kw = calibration.calibrate_scales(m32, [nc.synth_images(8, size).to(dev)])
    m = MobileNetV1_Q(3, 7).eval()
    m.load_state_dict(sd)
    nc.set_scales(m, ka, kw)
    plan = engine.compile_mobilenetv1(m.to(dev), batch, size)
    plan.input.copy_(nc.synth_images(batch, size).to(dev))
    for _ in range(2):
        plan.run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    plan.run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled", plan.launches_per_step, "launches")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--table":
        table(sys.argv[2])
    else:
        main()
