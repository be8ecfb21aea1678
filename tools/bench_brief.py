import json, sys
for line in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin):
    line = line.strip()
    if not line.startswith('{'):
        continue
    d = json.loads(line)
    print(json.dumps({"fps": round(d["value"]), "ms": round(d["ms_per_step"], 3), "kern": {k: round(v, 3) for k, v in d["kernel_ms"].items()},
                      "Gvisit/s(kernel)": round(d["gvoxel_visits_per_s_integrate_kernel"], 1), "frac": round(d["roofline"]["frac"], 3),
                      "e2e_fps": round(d["e2e"]["value"]), "e2e_ms": round(d["e2e"]["ms_per_step"], 2), "blocks": d["active_blocks"],
                      "upd": round(d["updated_voxel_fraction"], 3), "mesh": d["mesh"]}))
