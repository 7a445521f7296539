# A/B plan of the current session (sourced by gpu_session.sh)
one cur C3 X=1
one pr8 C3 X=1
one pr32 C3 X=1
one pr64 C3 X=1
one cur C4 X=1
one pr32 C4 X=1
one pr64 C4 X=1
one cur C5s X=1
one pr32 C5s X=1
cp ab/cur.so $LIB
