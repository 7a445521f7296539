# A/B plan of the current session (sourced by gpu_session.sh)
one split CM RTCUDA_NO_MATERIAL_SPLIT=1
one split CM X=1
one split CD RTCUDA_NO_MATERIAL_SPLIT=1
one split CD X=1
one split C3 X=1
cp ab/split.so $LIB
