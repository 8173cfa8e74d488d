"""profiles/launches_<tag>.csv (ncu launch list of ONE captured step) -> profiles/gemm_traffic_<tag>.json, the DRAM bytes of the
tf_gemm_kernel launches that bench.py quotes as `roofline.traffic` (with its source).  usage: python tools/gemm_traffic.py r2"""
import collections, csv, json, os, re, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(root, "profiles", f"launches_{tag}.csv")
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
byid = collections.OrderedDict()
for r in rows:
    byid.setdefault(r[0], {"name": r[4]})[r[12]] = float(r[14])
g = [d for d in byid.values() if "tf_gemm_kernel" in d["name"]]
rd, wr = sum(d.get("dram__bytes_read.sum", 0) for d in g), sum(d.get("dram__bytes_write.sum", 0) for d in g)
out = {"dram_bytes_per_step": rd + wr, "launches": len(g), "all_launches_of_the_step": len(byid),
       "source": f"ncu launch list profiles/launches_{tag}.csv of `bench.py --quick --steps 2 --warmup 3` (one captured step, "
                 f"dram__bytes_read.sum + dram__bytes_write.sum over its {len(g)} tf_gemm_kernel launches; ncu serialises the launches "
                 "with cold caches, so activations that stay L2-resident in the real step are counted as DRAM reads here). "
                 "Algorithmic: 1.72 GB of fp16 weights + 0.7 GB of activations."}
json.dump(out, open(os.path.join(root, "profiles", f"gemm_traffic_{tag}.json"), "w"), indent=1)
print(out)
