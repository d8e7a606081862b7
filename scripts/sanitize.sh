# compute-sanitizer passes over a small slice of the GPU suite (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards
# of the per-gene tables; synccheck: barrier misuse).  Slow (10-50x): keep the slice small.
set -o pipefail
T="tests/test_gpu_parity.py::test_gpu_calls_vs_reference_golden tests/test_gpu_parity.py::test_gpu_lfc_shrink_vs_reference_golden tests/test_gpu_edge_cases.py"
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 5 python -m pytest $T -x -q -m gpu -k "two_level_n24 or factorial_n30 or zero_genes or grid or optimizer or cooks or size_factors or shrink_arg" 2>&1 | tail -6
  echo "exit $?"
done
