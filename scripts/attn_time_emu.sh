for e in 0 1 2 3 4; do echo "EMU=$e"; ADAFACE_EXP_EMU=$e timeout 120 python scripts/attn_time.py 2>&1 | head -2; done
