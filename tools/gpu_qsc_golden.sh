set -x
mkdir -p gpurun_out/golden
python tools/make_golden.py --out gpurun_out/golden qsc64.qsc qsctilt64.qsc > gpurun_out/golden_qsc.log 2>&1
cp gpurun_out/golden/qsc*.npz gpurun_out/golden/qsc*.txt tests/golden/
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -m pytest tests/test_qsc.py -q > gpurun_out/pytest_qsc.log 2>&1; tail -15 gpurun_out/pytest_qsc.log
