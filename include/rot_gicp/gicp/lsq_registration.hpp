// lsq_registration.hpp — kept so `#include <rot_gicp/gicp/lsq_registration.hpp>` keeps resolving.
// In the B200 build the optimiser (GN / LM / SO3-LM / translation-LM step logic of the reference's
// LsqRegistration, lsq_registration_impl.hpp:55-335) lives behind the C-ABI in librolo_gicp_b200.so
// (rolo_b200/csrc/lsq_host.hpp); the class template in rot_vgicp.hpp exposes its public members.
#ifndef ROLO_B200_LSQ_REGISTRATION_HPP
#define ROLO_B200_LSQ_REGISTRATION_HPP
#include <rot_gicp/gicp/gicp_settings.hpp>
#endif
