"""Print the bench lines of an experiment directory side by side: tools/show_exp.py gpurun_out/exp2 [names...]"""
import glob, json, os, sys
d = sys.argv[1]
files = sorted(glob.glob(os.path.join(d, "*.json")))
for f in files:
    try:
        x = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(os.path.basename(f), "ERR", e)
        continue
    r = x["roofline"]
    pn = r.get("pair_node") or {}
    sh = {k[:6]: round(v, 3) for k, v in (r.get("share_of_step") or {}).items()}
    print("%-14s value %9.0f e2e %9.0f frac %.3f C=%-5s mttkrp %.4f pair %.4f leaf %.4f  %s" % (
        os.path.basename(f)[:-5], x["value"], x["e2e"]["value"], r["frac"], r.get("columns_on_this_gpu"),
        pn.get("mttkrp_dmma_ms_per_launch") or r["ms_per_launch"], pn.get("pair_contraction_ms_per_launch") or 0,
        pn.get("leaf_ms_per_launch") or 0, sh))
