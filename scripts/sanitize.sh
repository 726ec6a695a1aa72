#!/bin/bash
# compute-sanitizer over scripts/sanitize_driver.py, one tool per pass; summaries land in gpurun_out/sanitizer_<tool>.log
# (copy them to profiles/ after reading).  usage: scripts/sanitize.sh [tools...]   default: memcheck racecheck synccheck
mkdir -p gpurun_out
tools="${@:-memcheck racecheck synccheck}"
for tool in $tools; do
    extra=""
    [ "$tool" = memcheck ] && extra="--leak-check no"
    timeout 1500 compute-sanitizer --tool $tool $extra --print-limit 30 --error-exitcode 9 \
        --log-file gpurun_out/sanitizer_${tool}_full.log python scripts/sanitize_driver.py > gpurun_out/sanitizer_${tool}_run.log 2>&1
    rc=$?
    { echo "# compute-sanitizer --tool $tool $extra python scripts/sanitize_driver.py  -> exit code $rc";
      tail -3 gpurun_out/sanitizer_${tool}_run.log; grep -c "=========     at " gpurun_out/sanitizer_${tool}_full.log | sed 's/^/# error records: /';
      grep "ERROR SUMMARY\|RACECHECK SUMMARY\|Error:\|Warning:\|Race reported" gpurun_out/sanitizer_${tool}_full.log | sort | uniq -c | head -40; } > gpurun_out/sanitizer_${tool}.log
    cat gpurun_out/sanitizer_${tool}.log
done
