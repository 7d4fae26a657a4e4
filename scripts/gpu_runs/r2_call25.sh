#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 400 python scripts/epoch_cpu_time.py > $O/r25_epoch_cpu_time.jsonl 2> $O/r25_epoch_cpu_time.err; cat $O/r25_epoch_cpu_time.jsonl; tail -3 $O/r25_epoch_cpu_time.err
