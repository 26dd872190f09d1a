#!/bin/bash
out=gpurun_out
python -m pytest tests/test_pipeline_gpu.py -m gpu -q > $out/r2_pytest_pipe_a.log 2>&1; echo alone rc=$?; tail -2 $out/r2_pytest_pipe_a.log
python -m pytest tests/test_pipeline_gpu.py -m gpu -q > $out/r2_pytest_pipe_b.log 2>&1; echo alone2 rc=$?; tail -2 $out/r2_pytest_pipe_b.log
python -m pytest tests/test_mcts_gpu.py tests/test_chess_rules.py tests/test_leafq_gpu.py tests/test_pipeline_gpu.py -m gpu -q > $out/r2_pytest_pipe_c.log 2>&1; echo seq rc=$?; tail -2 $out/r2_pytest_pipe_c.log
