# A/B plan of the current session (sourced by gpu_session.sh)
one cur C3 X=1
one smem C3 X=1
one pl C3 X=1
one rcp C3 X=1
one cur C4 X=1
one smem C4 X=1
one pl C4 X=1
one rcp C4 X=1
one smem CM X=1
one smem C5s X=1
cp ab/smem.so $LIB
