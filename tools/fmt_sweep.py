import sys, json
for l in sys.stdin:
    try:
        r = json.loads(l)
    except Exception:
        print(l.strip()); continue
    print(r["m"], r["n"], r["L"], round(r["crt_frac"], 3), round(r["crt_inv_frac"], 3))
