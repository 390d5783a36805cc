for cfg in "16 50 84 256 1024 1 1 res" "8 200 336 64 256 1 1 res" "16 100 168 128 512 1 1 res" "16 25 42 512 2048 1 1 res" "16 100 168 512 128 1 1 mask" "16 100 168 256 256 3 1 none"; do
   python tools/bench_one.py $cfg fwd | sed "s|^|  $cfg : |"
done
