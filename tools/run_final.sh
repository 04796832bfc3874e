python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1
python bench.py > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err
python bench.py --impl reference > gpurun_out/bench_ref_v4.json 2> gpurun_out/bench_ref_v4.err
cat gpurun_out/pytest_final.log; tail -2 gpurun_out/smoke_final.log; cat gpurun_out/bench_v4.json; cat gpurun_out/bench_ref_v4.json
