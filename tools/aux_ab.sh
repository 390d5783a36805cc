for cfg in "16 50 84 256 1024 1 1 res" "16 50 84 256 1024 1 1 resmask" "16 100 168 128 512 1 1 res" "16 100 168 128 512 1 1 resmask" "16 25 42 512 2048 1 1 res" "16 25 42 512 2048 1 1 resmask"; do
   python tools/bench_one.py $cfg fwd | sed "s|^|  $cfg : |"
done
