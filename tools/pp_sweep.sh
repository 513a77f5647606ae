set -x
for pp in 0 1 2 3; do
  echo "=== RZK_PP=$pp"
  RZK_PP=$pp timeout 300 python tools/quick_time.py 2>&1 | grep -E "^(commit|open_|linear|sum|flags)" 
done
echo "=== tests under RZK_PP=2"
RZK_PP=2 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== cta_sync sweep (pp=0)"
for cs in 0 1 2 4 8; do
  echo "--- RZK_CTA_SYNC=$cs"
  RZK_CTA_SYNC=$cs timeout 300 python tools/quick_time.py 2>&1 | grep -E "^(commit|open_verify)"
done
