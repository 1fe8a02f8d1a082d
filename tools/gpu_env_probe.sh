#!/bin/bash
# environment facts of a GPU box that the host runtime depends on
echo "== nproc $(nproc)"; free -g | head -2; df -h /dev/shm | tail -1
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader
nvidia-smi topo -m 2>/dev/null | head -12
ulimit -l
