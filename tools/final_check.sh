#!/bin/bash
# Final 1-GPU check of the committed tree: GPU suite, smoke, the default bench line and the reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > $O/final_pytest.log 2>&1; echo "rc=$?" >> $O/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "rc=$?" >> $O/final_smoke.log
python bench.py --impl reference --steps 20 --warmup 3 > $O/final_ref.json 2> $O/final_ref.err
python bench.py > $O/final_line.json 2> $O/final_line.err
echo done
