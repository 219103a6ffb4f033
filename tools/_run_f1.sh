set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/r2_pytest_f1.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest_f1.log
timeout 600 python tools/bench_model_train.py --steps 30 > gpurun_out/r2_model_train.log 2> gpurun_out/r2_model_train.err; echo "rc=$?"; tail -c 1500 gpurun_out/r2_model_train.log; tail -5 gpurun_out/r2_model_train.err
