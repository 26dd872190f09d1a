for cfg in "512 800 3 4 8 1" "512 800 3 4 16 1" "512 800 3 8 8 1" "512 800 3 1 32 1" "512 800 3 4 8 0" "512 800 3 4 32 0"; do
  timeout 200 python scripts/selfplay_gpu.py $cfg
done
