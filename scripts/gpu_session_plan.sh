# A/B plan of the current session (sourced by gpu_session.sh)
one cur C3 X=1
one stream C3 X=1
one cur C4 X=1
one stream C4 X=1
one cur CM X=1
one stream CM X=1
one cur C5s X=1
one stream C5s X=1
cp ab/cur.so $LIB
