#!/bin/bash
# compute-sanitizer passes over the parity tests (logs go to gpurun_out/, summaries to profiles/)
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_two_pass_approx_equals_scoring_every_candidate tests/test_gpu_parity.py::test_exact_scores tests/test_gpu_parity.py::test_token_score_matrices_match_the_oracle tests/test_gpu_api.py::test_one_rank_communicator_runs_the_sharded_c_path tests/test_gpu_encode.py::test_kmeans_kernels_match_torch"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $T -x -q > gpurun_out/r02_sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.txt
# the three TMEM / mbarrier kernels: K1 v2 (every config), K5 v5 (Q = 64), encode_assign (+ the k-means variant)
R="tests/test_gpu_parity.py::test_centroid_scores[base] tests/test_gpu_parity.py::test_exact_scores[q64] tests/test_gpu_encode.py::test_kmeans_kernels_match_torch tests/test_gpu_encode.py::test_encode_matches_oracle"
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest $R -x -q > gpurun_out/r02_sanitizer_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.txt
timeout 1500 compute-sanitizer --tool synccheck python -m pytest $R "tests/test_gpu_parity.py::test_two_pass_approx_equals_scoring_every_candidate[base]" -x -q > gpurun_out/r02_sanitizer_synccheck.txt 2>&1
echo "synccheck rc=$?" >> gpurun_out/r02_sanitizer_synccheck.txt
for f in gpurun_out/r02_sanitizer_*.txt; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" $f | tail -5; done
