#!/bin/bash
# "Install" the reference for the --impl reference arm of bench.py.
# 1) the prescribed offline pip install (fails: the reference ships no setup.py / pyproject.toml, it is a script tree)
# 2) fallback: verbatim, unmodified copy of its src/ into baseline/_ref (git-ignored; travels to the GPU box)
set -u
cd "$(dirname "$0")/.."
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference \
  > baseline/_ref_pip.log 2>&1
echo "pip exit code: $? (see baseline/_ref_pip.log)"
mkdir -p baseline/_ref
rm -rf baseline/_ref/src
cp -r /root/reference/src baseline/_ref/src
cp /root/reference/LICENSE.md baseline/_ref/ 2>/dev/null || true
( cd /root/reference && find src -type f -name '*.py' -o -name '*.yml' | sort | xargs sha256sum ) > baseline/_ref/SHA256SUMS
echo "copied $(find baseline/_ref/src -name '*.py' | wc -l) python files to baseline/_ref/src"
