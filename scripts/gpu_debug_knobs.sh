for D in 0 1 2 4 5 6 7; do
  echo "=== SPXB_UMMA_DEBUG=$D"
  SPXB_UMMA_DEBUG=$D SPXB_UMMA_TRACE=1 python scripts/gpu_trace.py C3 C5 2>&1 | grep -E "==|convert loop end|all issued|exit|globaltimer"
done
