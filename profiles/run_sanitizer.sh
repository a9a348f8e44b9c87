#!/bin/bash
# compute-sanitizer over every kernel family (profiles/run_sanitizer.py): memcheck on all cases, racecheck / synccheck on the
# shared-memory-exchanging ones, initcheck on the workspace users.  Each tool run has its own timeout.
mkdir -p gpurun_out
out=gpurun_out/${1:-r2}_sanitizer.log
: > $out
echo "== plain run (no tool)" >> $out
timeout 600 python profiles/run_sanitizer.py >> $out 2>&1 || { echo "plain run failed" >> $out; tail -30 $out; exit 1; }
run() { # tool, timeout, cases...
  tool=$1; lim=$2; shift 2
  echo "== compute-sanitizer --tool $tool : $*" >> $out
  timeout $lim compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python profiles/run_sanitizer.py "$@" > gpurun_out/san_tmp.log 2>&1
  rc=$?
  grep -E "^ok |ERROR SUMMARY|=========.*(Invalid|Race|hazard|Uninitialized|Barrier|error)" gpurun_out/san_tmp.log | head -60 >> $out
  echo "rc=$rc" >> $out
  if [ $rc -ne 0 ]; then tail -40 gpurun_out/san_tmp.log >> $out; fi
}
run memcheck 900 small sp stab lims glw glw_stab glw_lims dyn stream
run racecheck 900 small sp stab lims glw dyn
run synccheck 600 small sp stab glw dyn
run initcheck 900 small sp glw glw_stab dyn stream
rm -f gpurun_out/san_tmp.log
cat $out
