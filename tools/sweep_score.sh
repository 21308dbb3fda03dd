# Scoring-kernel tuning sweep (run under gpurun): warps per CTA x bulk-copy staging on / off.
for w in 1 2 4 8; do for ns in 0 1; do
  if [ $ns = 0 ]; then export B2S_SCORE_STAGE=1; else unset B2S_SCORE_STAGE; fi
  echo "warps=$w nostage=$ns"; B2S_SCORE_WARPS=$w timeout 300 python tools/bench_ppo_loss.py 2>/dev/null | grep "k_score" | cut -c1-120
done; done
