# Regenerate every golden artefact that is recorded from the reference's own sources
# (build container only: needs /root/reference; nothing here runs on the GPU box).
#   bash tools/regen_golden.sh            # all
#   bash tools/regen_golden.sh --only jet # one family of tests/golden/reference_vectors
set -e
cd "$(dirname "$0")/.."
python -W ignore oracle/gen_reference_vectors.py "$@"
if [ $# -eq 0 ]; then
  python -W ignore oracle/gen_reference_aux_vectors.py
  python -W ignore oracle/gen_reference_api.py
fi
python -m pytest tests/test_reference_vectors.py tests/test_api_parity.py -q -m "not gpu"
