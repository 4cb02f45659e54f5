/* CppDebugger.hpp — logging shim (TEST INFRASTRUCTURE, not product code).
 *
 * The reference links the un-vendored Lut99/CppDebugger logging library
 * (reference CMakeLists.txt:18,74). It carries no arithmetic; to compile the
 * reference's CPU backend as the parity oracle (oracle/Makefile) this header
 * supplies the handful of names the reference sources use: silent logging
 * macros, and a Fatal exception raised by DLOG(fatal, ...), which is the
 * reference's only error channel (reference src/Main.cpp:305-308).
 */
#ifndef RT3_ORACLE_CPPDEBUGGER_SHIM_HPP
#define RT3_ORACLE_CPPDEBUGGER_SHIM_HPP

#include <stdexcept>
#include <string>
#include <vector>
#include <unordered_map>

namespace CppDebugger {
    enum class Severity { auxillary, info, warning, nonfatal, fatal, vulkan_warning, vulkan_error };

    namespace SeverityValues {
        static constexpr Severity auxillary = Severity::auxillary;
        static constexpr Severity info = Severity::info;
        static constexpr Severity warning = Severity::warning;
        static constexpr Severity nonfatal = Severity::nonfatal;
        static constexpr Severity fatal = Severity::fatal;
        static constexpr Severity vulkan_warning = Severity::vulkan_warning;
        static constexpr Severity vulkan_error = Severity::vulkan_error;
    }

    struct Fatal : public std::runtime_error {
        explicit Fatal(const std::string& what_arg) : std::runtime_error(what_arg) {}
    };

    inline void shim_log(Severity sev, const std::string& message) {
        if (sev == Severity::fatal) { throw Fatal(message); }
    }
}

#define DSTART(NAME)
#define DENTER(NAME)
#define DLEAVE
#define DRETURN return
#define DINDENT
#define DDEDENT
#define DMUTE(NAME)
#define DUNMUTE(NAME)
#define DLOG(SEVERITY, MESSAGE) ::CppDebugger::shim_log((SEVERITY), (MESSAGE))

#endif
