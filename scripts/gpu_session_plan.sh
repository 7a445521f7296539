# A/B plan of the current session (sourced by gpu_session.sh)
