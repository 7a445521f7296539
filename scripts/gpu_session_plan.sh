# A/B plan of the current session (sourced by gpu_session.sh)
one base C3 X=1
one l0 C3 X=1
one l0 C3 RTCUDA_NO_LIGHT0=1
one l0sb2 C3 X=1
one base C4 X=1
one l0 C4 X=1
one base C5s X=1
one l0 C5s X=1
one base CM X=1
one l0 CM X=1
cp ab/l0.so $LIB
