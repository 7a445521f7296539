# A/B plan of the current session (sourced by gpu_session.sh)
one nosmem C3 X=1
one new C3 X=1
one sb2 C3 X=1
one nosmem C4 X=1
one new C4 X=1
one new CM X=1
one new C2 X=1
cp ab/new.so $LIB
