#!/bin/bash
# which stream goes first when the chain behind stage 1 is launched with PDL (it then needs less of the machine):
# B200_STREAM_PRIO = "tail,main,fft" (0 = lowest) x B200_PDL
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "2:2,1,0" "1:2,1,0" "1:1,1,2" "1:0,1,2" "1:1,2,0" "1:0,0,0" "1:1,0,2" "2:1,1,2" "2:0,0,0" "1:2,0,1"; do
  IFS=: read pdl prio <<< "$cfg"
  B200_PDL=$pdl B200_STREAM_PRIO=$prio timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu --c3 0 --c4 0 > gpurun_out/pp_tmp.json 2> gpurun_out/pp_tmp.err
  python - "$pdl" "$prio" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/pp_tmp.json').read().strip().splitlines()[-1])
    cs=d["config"]["chunk_sweep"]
    print("pdl=%s prio=%s value=%.0f" % (sys.argv[1], sys.argv[2], d["value"]), {k: round(x["value"]) for k, x in cs.items()}, [(g["group"], round(g["avg_ms"]*1e3, 1)) for g in d["roofline"]["by_group"]])
except Exception as e:
    print(sys.argv[1:], 'failed', e)
PY
done
