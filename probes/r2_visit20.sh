#!/bin/bash
for cfg in "1,1,2;2;2,1,1" "1;1;1" "1,1;3;2,1,1" "2;2;2"; do
  echo "chunks $cfg"; NLA_STREAM_CHUNKS="$cfg" timeout 120 python probes/time_host_opts2.py 2>&1 | head -1
done | tee gpurun_out/r2_stream_chunks2.txt
