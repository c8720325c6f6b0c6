LC3D_LIB=$PWD/liblc3d_H.so python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in G H G H; do
echo "variant $v"; LC3D_LIB=$PWD/liblc3d_$v.so python scripts/dev_profile_icp.py 1 8 2>&1 | tail -1
done
echo "variant p2p H"; LC3D_LIB=$PWD/liblc3d_H.so python scripts/dev_profile_icp.py 0 4 2>&1 | tail -1
LC3D_LIB=$PWD/liblc3d_H.so LC3D_STATS=1 python scripts/dev_profile_icp.py 1 2 2>&1 | grep "fitness\|warp \|cycles" | grep -v "^\[lc3d stats\] it" | tail -20
