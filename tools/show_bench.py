import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "value %.3e ms %.4f warm %.3e frac %.3f e2e %.3e"%(d["value"], d["ms_per_step"], d["value_l2_warm"], d["roofline"]["frac"], d["e2e"]["value"]), d["step_ms_percentiles"])
    for s in d.get("sweep",[]): print("    envs %8d ms %.4f value %.3e warm %.3e frac %.3f"%(s["envs"], s["ms_per_step"], s["value"], s["value_l2_warm"], s["frac"]))
