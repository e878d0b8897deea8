# compute-sanitizer over every kernel family of the product library (tools/sanitize_cases.py); logs -> gpurun_out/
mkdir -p gpurun_out
python tools/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/sanitize_plain.log
for tool in memcheck synccheck racecheck; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|failures" gpurun_out/sanitize_$tool.log | tail -3
done
