import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print("headline", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("timed_passes"))
print("fit ms", d["roofline_fit"]["avg_launch_ms"], "fused", d["roofline"]["avg_launch_ms"], d["roofline"]["frac"])
print("all-sur", d.get("predict_all_surrogates"))
print("cpu", d.get("cpu_baseline"))
print("clocks", d.get("clocks"))
for k,v in d.get("sub",{}).items():
    if v is None: continue
    if k=="predict_sweep":
        print(k, v["best_frac_of_dmma_peak"], v.get("cpu_baseline"), v["record_wall_s"])
        for r in v["rows"]: print("   ", r["d"], r["n"], r["N"], round(r["ms"],3), "%.3e"%r["candidates_per_s"], round(r["frac_of_dmma_peak"],3))
    else:
        print(k, v.get("value"), v.get("ms_per_step"), "e2e", (v.get("e2e") or {}).get("value"), "cpu", (v.get("cpu_baseline") or {}).get("value"), "wall", v.get("record_wall_s"), v.get("predict_all_surrogates"), v.get("per_rank_ms_per_step"))
