for cfg in "split_count=0" "split_chunks=1" "split_chunks=2"; do echo "== bpe $cfg"; timeout 200 python profiles/scripts/launches.py 1000000000 2 bpe timing=1 $cfg 2>&1 | grep timing | tail -1; done
