timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for c in 1 16 64 256; do timeout 200 ./sdr-server_b200/bin/dropin_bench $c 50 2>&1 | tail -1; done
echo LANES=1; for c in 1 64 256; do XLATING_B200_LANES=1 timeout 200 ./sdr-server_b200/bin/dropin_bench $c 50 2>&1 | tail -1; done
echo LANES=4; for c in 64 256; do XLATING_B200_LANES=4 timeout 200 ./sdr-server_b200/bin/dropin_bench $c 50 2>&1 | tail -1; done
