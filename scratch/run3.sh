timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for c in 1 16 64 256; do timeout 100 ./sdr-server_b200/bin/dropin_bench $c 50 2>&1 | tail -1; done
timeout 100 ./sdr-server_b200/bin/dropin_bench 512 30 2>&1 | tail -1
echo SHARE=0; for c in 64 256; do XLATING_B200_SHARE=0 timeout 100 ./sdr-server_b200/bin/dropin_bench $c 50 2>&1 | tail -1; done
