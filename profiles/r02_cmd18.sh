mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_fulltests.txt
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02_smoke.txt 2>&1
echo "smoke rc=$?" >> gpurun_out/r02_smoke.txt
