// oracle/ref_shim/kdl_parser/kdl_parser.hpp — interface only: URDF parsing is outside the compiled-reference harness (the
// driver builds the KDL::Tree from the segment table instead).  TEST INFRASTRUCTURE.
#ifndef STOMP_REF_SHIM_KDL_PARSER
#define STOMP_REF_SHIM_KDL_PARSER
#include <string>
#include <kdl/tree.hpp>
namespace kdl_parser { inline bool treeFromString(const std::string&, KDL::Tree&) { return false; } }
#endif
