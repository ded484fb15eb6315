timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "^E  |passed|failed|Error" | head -12 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
