#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [gpurun options] -- 'command'
# Retries while gpurun answers "busy" (exit 3: nothing charged), every 90 s, up to 40 times.
log="$1"; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 90
done
exit 3
