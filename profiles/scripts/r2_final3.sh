bash profiles/scripts/r2_final2.sh
bash profiles/scripts/r2_evidence2.sh 2>&1 | tail -24
