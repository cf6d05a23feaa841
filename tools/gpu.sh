#!/bin/bash
# build every native artefact (stale-checked), then run the given command on the B200 box
set -e
cd /root/repo
python __graft_entry__.py > /tmp/build.log 2>&1 || { tail -30 /tmp/build.log; exit 1; }
T=${GPU_TIMEOUT:-900}
exec /usr/local/graft/bin/gpurun --timeout $T -- "$@"
