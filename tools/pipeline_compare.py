import json,glob,sys
tag=sys.argv[1]
a=json.load(open(f"gpurun_out/pipeline_{tag}_w1_r0.json"))
b={}
for f in glob.glob(f"gpurun_out/pipeline_{tag}_w2_r*.json"): b.update(json.load(open(f))["refs"])
print(tag, "stats", [(s["name"], round(s["device_ms"],1), round(s["exchange_ms"],2)) for s in a["stats"]])
print(tag, "bit-identical 1 vs 2 GPUs:", all(a["refs"][k]["sha"]==b[k]["sha"] for k in a["refs"]), "acc ref4", a["refs"]["4"]["acc"], "mean cost", round(a["refs"]["4"]["mean_cost"],4))
