bash tools/profile_round.sh r02_v12 c3 2>&1 | tail -60
