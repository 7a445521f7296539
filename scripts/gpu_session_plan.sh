# A/B plan of the current session (sourced by gpu_session.sh)
one cur C3 X=1
one qpf2 C3 X=1
one cur CM X=1
one qpf2 CM X=1
cp ab/cur.so $LIB
