#!/bin/bash
# usage: scratch/ab.sh "ENV_A" "ENV_B" [reps]  -- alternates bench runs with the two environments
A="$1"; B="$2"; R="${3:-2}"
fmt='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(sys.argv[1], "step %.1f fwd %.1f bwd %.1f e2e %.1f us" % (1e3*d["ms_per_step"], 1e3*r["fwd_ms"], 1e3*r["bwd_ms"], 1e3*d["e2e"]["ms_per_step"]))'
for i in $(seq $R); do
  env $A python bench.py --steps 100 --warmup 10 --no-cpu --no-extra 2>/dev/null | python -c "$fmt" "A[$A]"
  env $B python bench.py --steps 100 --warmup 10 --no-cpu --no-extra 2>/dev/null | python -c "$fmt" "B[$B]"
done
