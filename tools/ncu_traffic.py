"""DRAM traffic of the wide tcgen05 GEMMs (encoder + cross-K/V projection) per launch, measured by ncu, next to their
algorithmic operand bytes -> profiles/r2_traffic.json (read by bench.py for `roofline.traffic`).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
        -k regex:"gemm_tc2_kernel|gemm_tc_kernel" --csv --log-file gpurun_out/traffic.csv \
        python tools/prof_step.py --enc-layers 4 --max-length 5
    python tools/ncu_traffic.py gpurun_out/traffic.csv 4 64 [out.json]"""
import csv, json, os, re, sys
from collections import defaultdict

path, enc_layers, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_traffic.json")
lines = [l for l in open(path) if not l.startswith("==")]
per = defaultdict(lambda: defaultdict(float))   # launch id -> metric -> bytes
names = {}
for r in csv.DictReader(lines):
    m = r.get("Metric Name", "")
    if not m.startswith("dram__bytes"):
        continue
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
    per[int(r["ID"])][m] += v
    names[int(r["ID"])] = re.sub(r"\(.*", "", r["Kernel Name"])
ids = [i for i, n in names.items() if "gemm_tc2_kernel" in n or n.strip().endswith("gemm_tc_kernel")]
rd = sum(per[i]["dram__bytes_read.sum"] for i in ids)
wr = sum(per[i]["dram__bytes_write.sum"] for i in ids)
# algorithmic operand bytes of the same launches (kotoba architecture: d 1280, ffn 5120, 128 mels, 2 decoder layers, bf16)
d, F, S, T2, mel3, L_dec = 1280, 5120, 1500, 3000, 384, 2
M = B * S
gemm = lambda m, n, k, out_b, resid=0: m * k * 2 + n * k * 2 + m * n * out_b + resid
alg = gemm(B * T2, d, mel3, 2) + gemm(M, d, 3 * d, 4)                                   # conv1, conv2 (+pos table, negligible)
alg += enc_layers * (gemm(M, 3 * d, d, 2) + gemm(M, d, d, 4, M * d * 4) + gemm(M, F, d, 2) + gemm(M, d, F, 4, M * d * 4))
alg += L_dec * gemm(M, 2 * d, d, 2)                                                     # cross-K/V projections
n_expected = 2 + 4 * enc_layers + L_dec
res = {"source": os.path.basename(path), "encoder_layers_captured": enc_layers, "batch": B,
       "encoder_gemm_launches": len(ids), "expected_launches": n_expected,
       "encoder_gemm_dram_bytes": rd + wr, "encoder_gemm_dram_read_bytes": rd, "encoder_gemm_dram_write_bytes": wr,
       "encoder_gemm_algorithmic_bytes": alg, "traffic_over_algorithmic": (rd + wr) / alg if alg else None,
       "per_kernel": {}}
agg = defaultdict(lambda: [0, 0.0, 0.0])
for i in ids:
    a = agg[names[i]]
    a[0] += 1; a[1] += per[i]["dram__bytes_read.sum"]; a[2] += per[i]["dram__bytes_write.sum"]
for n, (c, r_, w_) in agg.items():
    res["per_kernel"][n] = {"launches": c, "read_MB_per_launch": r_ / c / 1e6, "write_MB_per_launch": w_ / c / 1e6}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
