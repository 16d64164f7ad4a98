"""profiles/traffic_cfg3.json from an ncu DRAM-traffic CSV of one or more cfg-3 steps
(scripts/final_n1.sh: --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none).
Usage: python scripts/make_traffic_json.py <csv> "<library version string>" """
import csv, json, sys

NAMES = [("prep_kernel", "prep_norms"), ("umma_kernel<0", "fwd_logits_stats"), ("combine_kernel", "combine_stats"),
         ("umma_kernel<1", "bwd_recompute_g"), ("umma_kernel<2", "dw_gemm"), ("umma_kernel<3", "dx_gemm"),
         ("dx_finish_kernel", "dx_finish")]
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, cnt = {}, {}
for r in rows[1:]:
    if not r[mi].startswith("dram__bytes"):
        continue
    for pat, nm in NAMES:
        if pat in r[ki]:
            tot[nm] = tot.get(nm, 0.0) + float(r[vi].replace(",", "")) * scale[r[ui]]
            cnt[nm] = cnt.get(nm, 0) + 1
steps = max(1, min(cnt.values()) // 2)             # two metrics per launch
kern = {nm: int(tot[nm] / (cnt[nm] / 2)) for _, nm in NAMES if nm in tot}
out = {"library_version": sys.argv[2], "workload": "cfg3",
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none (single pass, L2 state "
                 "carried from the preceding kernel), kernels serialised by ncu; mean of %d steps" % steps,
       "kernels": kern, "step_total": int(sum(kern.values()))}
json.dump(out, open("profiles/traffic_cfg3.json", "w"), indent=1)
print(json.dumps(out))
