import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(sys.argv[1] if len(sys.argv)>1 else "", d["config"], d.get("generator",""), d.get("density",""), d.get("query",""), "ms=%.4f frac=%.3f gbs=%.0f"%(d["ms"], d["frac"], d["achieved_gbs"]))
