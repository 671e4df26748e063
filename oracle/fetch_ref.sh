#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Makes the *unmodified* reference importable on the GPU box.
#
# The reference is pure Python, so "compiling it from the sources where they lie" is a copy:
# this recipe mirrors the packages the hot-path tests and the CPU baseline import from /root/reference into
# oracle/_ref/ (git-ignored, NOT gpurun-ignored -> it travels with the snapshot exactly like the built .so).
# Nothing under oracle/_ref/ is ever committed, edited, or imported by the product (betty_b200/).
#   betty/                                    the library: hypergradient/{neumann,cg,darts,sama}.py, Engine, problems
#   test/test_regression.py                   the reference's own end-to-end regression of this path (loss < 0.48)
#   examples/neural_architecture_search/*.py  Network(16,10,8)/Architecture: config 4's model, to pin the restatement
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$SRC/betty" ]; then
  echo "fetch_ref: $SRC/betty not found (GPU box: using the prebuilt oracle/_ref)"; exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/test" "$DST/examples/neural_architecture_search"
cp -r "$SRC/betty" "$DST/betty"
cp "$SRC/test/__init__.py" "$SRC/test/test_regression.py" "$DST/test/"
for f in model_search.py operations.py genotypes.py; do
  cp "$SRC/examples/neural_architecture_search/$f" "$DST/examples/neural_architecture_search/"
done
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$DST/REVISION"
echo "fetch_ref: mirrored $(find "$DST" -name '*.py' | wc -l) reference files into $DST"
