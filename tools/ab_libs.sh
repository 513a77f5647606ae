# A/B of several builds of the library in ONE gpurun call (same box, same clocks), rounds interleaved:
# usage: [ROUNDS=3] bash tools/ab_libs.sh "<phase> [<phase> ...]" <lib.so> [<lib.so> ...]
PHASES=$1; shift
for round in $(seq 1 ${ROUNDS:-3}); do
  for lib in "$@"; do
    for p in $PHASES; do
      RZK_LIB_PATH=$PWD/$lib python tools/ab_time.py $p -- "" 2>&1 | grep M/s | sed "s|default  *|$(basename $lib) round $round|"
    done
  done
done
