#!/bin/bash
# stream experiment: device-resident value of the bench workload under different stream priorities (tail,main,fft : main
# stream priority of the bench) and with stage 1 serialised behind the spectrum branch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "2,1,0:-1:0" "2,1,0:-1:1" "2,0,0:0:1" "0,0,0:0:1" "2,1,1:-1:1"; do
    IFS=: read pr mp fs <<< "$cfg"
    B200_FFT_SERIAL="$fs" B200_STREAM_PRIO="$pr" timeout 300 python bench.py --quick --steps 6 --warmup 3 --no-cpu --c3 0 --c4 0 --main-prio "$mp" > gpurun_out/prio_tmp.json 2> gpurun_out/prio_tmp.err
    python - "$pr" "$mp" "$fs" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/prio_tmp.json').read().strip().splitlines()[-1])
    g={x['group']:round(x['avg_ms']*1e3,1) for x in d['roofline']['by_group']}
    print(sys.argv[1:], round(d['value']), g)
except Exception as e:
    print(sys.argv[1:], 'failed', e)
PY
done
