# Runs the UNMODIFIED reference package (installed under baseline/_ref, never copied into the repo) and its own
# test files on the B200 backend through the MKL-symbol shim.  Usage: bash scripts/run_reference_suite.sh [files...]
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=$ROOT/baseline/_ref
if [ ! -d "$REF/sparse_dot_mkl" ]; then echo "baseline/_ref/sparse_dot_mkl is not installed"; exit 2; fi
export MKL_RT=$ROOT/sparse_dot_b200/libsdb200_mkl.so
export PYTHONPATH=$REF
cd "$REF"
python -c "import sparse_dot_mkl as m; print('reference', m.__version__, 'on:', m.get_version_string()); import sparse_dot_mkl._mkl_interface as i; print('MKL_INT', i.mkl_interface_integer_dtype())"
FILES=${*:-"test_mkl.py test_sparse_dense.py test_sparse_sparse.py test_sparse_vector.py test_gram_matrix.py"}
for f in $FILES; do
  echo "=== $f"
  python -m pytest sparse_dot_mkl/tests/$f -q --no-header -p no:cacheprovider 2>&1 | grep -E "passed|failed|^FAILED|^ERROR" | tail -40
done
