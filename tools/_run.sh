echo "vote=0"; LPC_VOTE=0 python tools/win_probe.py 0
echo "vote=1"; LPC_VOTE=1 python tools/win_probe.py 0
