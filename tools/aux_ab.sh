for v in 0 1; do
 echo "UT2_AUX_DBL=$v"
 for cfg in "16 50 84 256 1024 1 1 res" "8 200 336 64 256 1 1 res" "16 100 168 128 512 1 1 res" "16 25 42 512 2048 1 1 res" "16 100 168 512 128 1 1 mask" "16 50 84 1024 256 1 1 mask"; do
   UT2_AUX_DBL=$v python tools/bench_one.py $cfg fwd | sed "s|^|  $cfg : |"
 done
done
