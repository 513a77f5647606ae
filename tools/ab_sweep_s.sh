# Phase-mixing / lowering sweep of the programs that moved to the 26-bit primes (one gpurun call, device resident, CUDA events)
python tools/ab_time.py commit -- "" "RZK_TUNE=commit_pp=0" "RZK_TUNE=commit_pp=21" "RZK_TUNE=commit_pp=22" "RZK_TUNE=commit_small=0"
python tools/ab_time.py sum_commit -- "" "RZK_TUNE=mulsum2_pp=0" "RZK_TUNE=mulsum2_pp=21" "RZK_TUNE=mulsum2_pp=22" "RZK_TUNE=mulsum_small=0" "RZK_TUNE=wave_fit=0"
python tools/ab_time.py sum_verify -- "" "RZK_TUNE=verify_w_pp=0" "RZK_TUNE=verify_w_pp=22" "RZK_TUNE=verify_w_pp=2" "RZK_TUNE=mulsum_small=0"
python tools/ab_time.py linear_commit -- "" "RZK_TUNE=wave_fit=0" "RZK_TUNE=mulsum_small=0"
python tools/ab_time.py linear_verify -- "" "RZK_TUNE=wave_fit=0" "RZK_TUNE=verify_w_pp=22"
python tools/ab_time.py open_verify -- "" "RZK_TUNE=verify_pp=21" "RZK_TUNE=verify_pp=0" "RZK_TUNE=verify_pp=2"
