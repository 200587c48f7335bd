# chunk schedule of kmu_sketch_pmh3a_host: growth per chunk (percent) and first chunk = total / div
for cfg in "50 128" "25 64" "30 64" "25 48" "20 64" "25 96" "50 128" "25 64"; do set -- $cfg; echo "grow=$1 first_div=$2"; KMU_HOST_GROW_PCT=$1 KMU_HOST_FIRST_DIV=$2 CALLS=7 python scripts/prof_e2e.py | head -1; done
