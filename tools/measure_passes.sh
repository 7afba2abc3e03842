# Per-pass helper measurements (one B200; run under gpurun from the repo root): full -m gpu test run, canary / determinism
# checker, tools/bench_passes.py for the three robots, ncu captures of the body-per-lane gradient fpass.
set -u
O=gpurun_out
cap() {
  local k=$1 stem=$2; shift 2
  ncu --set full --clock-control none --import-source on -k $k -c 1 -o $O/$stem "$@"
  python tools/ncu_summary.py $O/$stem.ncu-rep $O/summ_$stem.txt > /dev/null 2>&1
  python tools/ncu_lines.py $O/$stem.ncu-rep 2>/dev/null | awk '{ if ($4+0 >= 0.8 || $6+0 >= 0.8) print }' | cut -c1-260 >> $O/summ_$stem.txt
  rm -f $O/$stem.ncu-rep
}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/sanitize.py > $O/r02_sanitize.json 2> $O/r02_sanitize.err; tail -c 200 $O/r02_sanitize.json; tail -2 $O/r02_sanitize.err
: > $O/r02_passes.jsonl
python tools/bench_passes.py --reps 5 --batch 1048576 >> $O/r02_passes.jsonl 2>/dev/null
python tools/bench_passes.py --reps 5 --robot hyq --batch 262144 >> $O/r02_passes.jsonl 2>/dev/null
python tools/bench_passes.py --reps 5 --robot atlas --batch 65536 >> $O/r02_passes.jsonl 2>/dev/null
if [ "${1:-}" = "ncu" ]; then
cap regex:grad_fpass_level r02_prof_grad_fpass_level_iiwa14_f64 python tools/bench_passes.py --reps 1 --batch 1048576 > $O/r02_ncu5.log 2>&1
cap regex:grad_fpass_level r02_prof_grad_fpass_level_atlas_f64 python tools/bench_passes.py --reps 1 --robot atlas --batch 65536 > $O/r02_ncu5b.log 2>&1
fi
grep -c pass $O/r02_passes.jsonl
