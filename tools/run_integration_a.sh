#!/bin/bash
# Stage the UNMODIFIED reference main.py where the GPU box can see it (oracle/_ref/ is git-ignored: never committed), run
# tools/integration_a.py there, remove the staged copy again.
set -e
cd "$(dirname "$0")/.."
mkdir -p oracle/_ref && cp /root/reference/main.py oracle/_ref/main.py
/usr/local/graft/bin/gpurun --timeout 600 -- 'python tools/integration_a.py 2>&1 | grep -v "Warning\|warn" | tail -60'
rm -f oracle/_ref/main.py; rmdir oracle/_ref 2>/dev/null || true
