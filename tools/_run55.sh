cd /root/repo
UP3D_SIDE_INLINE=1 timeout 300 python tools/la_probe.py 2>&1 | grep "flush\|Error" | head -3
