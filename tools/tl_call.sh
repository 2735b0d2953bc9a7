#!/bin/bash
mkdir -p gpurun_out
NO_TESTS=1 bash tools/ab_call.sh tools/ab_cfg.txt ab14
