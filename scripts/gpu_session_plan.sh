# A/B plan of the current session (sourced by gpu_session.sh)
one cur C3 X=1
one b8 C3 X=1
one r8 C3 X=1
one r12 C3 X=1
one t10 C3 X=1
one t14 C3 X=1
one b8 C4 X=1
cp ab/cur.so $LIB
