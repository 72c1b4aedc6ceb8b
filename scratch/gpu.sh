#!/bin/bash
# build (fail loudly) + import check, then gpurun "$@"
set -e
cd /root/repo
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { tail -20 /tmp/build.log; echo BUILD FAILED; exit 1; }
python -c "import sys; sys.path.insert(0,'/root/repo'); import ffb200.train, ffb200.ops, ffb200.models.FactorFields, bench" || { echo IMPORT FAILED; exit 1; }
exec /usr/local/graft/bin/gpurun "$@"
