for cfg in "4 8" "8 8" "16 8" "8 4" "8 16" "16 32" "2 8"; do set -- $cfg; IID_DL_THREADS=$1 IID_DL_CHUNK_MB=$2 python scripts/gpu_e2e_time.py 2>&1 | tail -1; done
nproc
