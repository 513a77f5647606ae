set -x
RZK_LIB_PATH=$PWD/ring-zk_b200/_build/libringzk_b200_light.so timeout 300 python tools/quick_time.py 2>&1 | grep -E "^(open_respond|linear_respond|sum_respond|.*Error)"
