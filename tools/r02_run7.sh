#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
( time python -m pytest tests -m gpu -q -s --maxfail=12 ) > $O/r02f_pytest.log 2>&1
tail -8 $O/r02f_pytest.log | cut -c1-300
grep -h "multiscatter\|closures_multi\|config1" $O/r02f_pytest.log | grep rmse | cut -c1-260
python bench.py --steps 3 --warmup 3 --configs cube,cornell > $O/r02f_bench.json 2> $O/r02f_bench.err
tail -c 400 $O/r02f_bench.err
