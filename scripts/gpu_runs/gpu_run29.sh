mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "peer" > gpurun_out/pytest_peer.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_peer.log | cut -c1-300
