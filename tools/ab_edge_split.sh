#!/bin/bash
# A/B of the graded chunk sizes of the host pipeline (host_io.cu: the exposed first / last chunk cut into 8 + 8 + rest fields):
# plugin-path round trip at cfg3 (tools/e2e_profile.py), SHB200_HOST_EDGE_SPLIT=0 / 1, two runs each.
for S in 0 1 0 1; do
  echo "== SHB200_HOST_EDGE_SPLIT=$S"
  SHB200_HOST_EDGE_SPLIT=$S python tools/e2e_profile.py 2>&1 | head -1
done
