#!/bin/bash
# tools/ab.sh NAME... — bench (device-resident leg only) every build in bronko_b200/csrc/variants/ named on the command
# line, one line per build: ms per sample with four samples in flight, then the single-sample stage times.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# NAME may carry ":stages" (BK_ABLATE of a -DBK_ABLATE build) and "@S" (samples in flight), e.g. abl:map,noise@2
for spec in "$@"; do
  v=${spec%%[:@]*}; abl=""; infl=4
  [[ $spec == *:* ]] && { abl=${spec#*:}; abl=${abl%%@*}; }
  [[ $spec == *@* ]] && infl=${spec##*@}
  BK_ABLATE=$abl BRONKO_B200_LIB=$PWD/bronko_b200/csrc/variants/$v.so python bench.py --no-e2e --no-cpu-baseline --no-fastq --no-sharded --in-flight $infl --depth ${DEPTH:-10000} --steps ${STEPS:-80} > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" "$spec" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/ab_%s.json" % v).read().strip().splitlines()[-1])
    a = d["stage_ms_single_sample"]
    print("%-28s ms/step %.4f  latency %.3f | scan %.3f left %.3f fin %.3f map %.3f score %.3f | clocks %s nvar %d" % (
        sys.argv[2], d["ms_per_step"], d["latency_ms_single_sample"], a["scan_ms"], a["leftover_ms"], a["finalize_ms"], a["map_ms"],
        a["score_ms"], d["clocks"]["sm_mhz"], d["result_check"]["n_variants"]))
except Exception as e:
    print(v, "FAILED", e, open("gpurun_out/ab_%s.err" % v).read()[-500:])
PY
done
