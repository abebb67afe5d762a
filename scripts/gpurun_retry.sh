#!/bin/bash
# retry `gpurun` while the pod answers "busy" (exit 3: nothing charged); any other status is final
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
