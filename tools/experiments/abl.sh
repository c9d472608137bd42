for ab in 0 1 2 4 6 7; do echo "== decode kernel ablate=$ab"; python tools/microbench.py --M 1 --shapes llama8b --ablate $ab --reps 5 2>&1 | grep "N="; done
