# A/B plan of the current session (sourced by gpu_session.sh)
one cur C3 X=1
one nochunk C3 X=1
one fetch C3 X=1
one chunk32 C3 X=1
one chunk128 C3 X=1
one cur C4 X=1
one fetch C4 X=1
one cur C5s X=1
one fetch C5s X=1
one cur C2 X=1
one fetch C2 X=1
cp ab/fetch.so $LIB
