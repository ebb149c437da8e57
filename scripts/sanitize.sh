#!/bin/sh
# compute-sanitizer evidence: memcheck, racecheck and initcheck over the forward and
# backward kernels at the level-2 and level-6 shapes, both configurations.  The level-2 shape runs at its
# full batch (1344 tiles on 148 persistent CTAs: ~9 tiles per CTA, the regime in which the mbarrier / LDS
# ordering of the rings matters).  usage (GPU box, repo root): sh scripts/sanitize.sh r02
R=${1:-r02}
O=gpurun_out
mkdir -p $O
SUM=$O/${R}_sanitizer_summary.txt
: > $SUM
run() {   # tool, tag, prof_case args...
    tool=$1; tag=$2; shift 2
    log=$O/${R}_sanitizer_${tool}_${tag}.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 \
        python scripts/prof_case.py "$@" > $log 2>&1
    rc=$?
    echo "== $tool $tag (prof_case $*): exit $rc" >> $SUM
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|last iter" $log | head -12 >> $SUM
}
for tool in memcheck racecheck; do
    run $tool level2_canon fwdbwd level2 iid canon 1
    run $tool level2_refcfg fwdbwd level2 iid ref 1
    run $tool level6_canon fwdbwd level6 iid canon 1
    run $tool level6_refcfg fwdbwd level6 iid ref 1
    run $tool level3_canon fwdbwd 8,64,48,56 iid canon 1
done
run initcheck level2_canon fwdbwd level2 iid canon 1
run initcheck level6_canon fwdbwd level6 iid canon 1
cat $SUM
