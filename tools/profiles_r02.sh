#!/bin/bash
# turns the reports tools/gpu_profile.sh brought back in gpurun_out/ into the tracked summaries under profiles/ (runs without a GPU)
cd "$(dirname "$0")/.."
python tools/ncu_summary.py gpurun_out/r02_xd_tma.ncu-rep > profiles/r02_ncu_full_xd_tma.txt
python tools/ncu_summary.py gpurun_out/r02_tails_reg.ncu-rep > profiles/r02_ncu_full_tails_reg.txt
python tools/ncu_summary.py gpurun_out/r02_fftr.ncu-rep > profiles/r02_ncu_full_fft_register.txt
cp gpurun_out/r02_launches_bench.csv profiles/r02_launches_bench.csv
cp gpurun_out/r02_trace.txt profiles/r02_trace_pipelined.txt
python tools/ncu_traffic.py gpurun_out/r02_xd_tma.ncu-rep 16777216 profiles/r02_traffic.json
sed -i 's#profiles/r02_xd_tma.txt#profiles/r02_ncu_full_xd_tma.txt#' profiles/r02_traffic.json
# SASS of the shipped library: the hot loop of the dominant kernel and the proof of the TMA / mbarrier path
cuobjdump -sass -fun '_Z8k_xd_tmaILi5ELi5ELi10ELi256ELi2EEv8XdParams6XtGeom14CUtensorMap_st' sdrplusplus_b200/libb200dsp.so > /tmp/xt.sass
{ echo "# cuobjdump -sass of k_xd_tma<5,5,10,256,2> in the shipped libb200dsp.so (sm_100a): instruction census and the start of the unrolled filter loop";
  echo "# UTMALDG = cp.async.bulk.tensor (TMA), SYNCS = mbarrier, FFMA2 with a UR operand = packed FMA whose tap comes from the constant bank";
  for m in UTMALDG SYNCS FFMA2 LDS.128 LDS.64 LDGSTS ULDC STG; do echo "$m $(grep -c "$m" /tmp/xt.sass)"; done;
  echo "# ---- first 60 instructions of the filter loop (from the first LDS.128)";
  awk '/LDS.128/{f=1} f{print} ' /tmp/xt.sass | grep -v "^\s*/\* 0x" | head -60; } > profiles/r02_sass_xd_tma.txt
