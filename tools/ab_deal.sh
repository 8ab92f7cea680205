for m in dyn static dyn static; do
  if [ $m = static ]; then export PDE_B200_STATIC_DEAL=1; else unset PDE_B200_STATIC_DEAL; fi
  MODE=$m python tools/ab_deal.py
done
