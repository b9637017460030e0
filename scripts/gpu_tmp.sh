cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q -x -s -k "precision or tc or hier or golden or human" 2>&1 | grep -v "^$" | tail -30
