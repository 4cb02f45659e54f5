/* CppDebugger.hpp — the error channel of the host backend.
 *
 * The reference reports every failure through DLOG(fatal, msg), which throws
 * CppDebugger::Fatal; main() catches it and exits with -1 (reference
 * src/Main.cpp:305-308, e.g. SequentialRenderer.cpp:241,249). The backend keeps
 * that behaviour. Standalone, this header provides those names; inside the
 * reference tree the real Lut99/CppDebugger header is found first.
 *
 * The same header lets oracle/Makefile compile the reference's own CPU sources
 * (which link the un-vendored Lut99/CppDebugger library, reference
 * CMakeLists.txt:18,74) as the parity checker: it carries no arithmetic, only
 * the handful of names those sources use. -DRT3_DEBUGGER_SILENT drops the
 * warnings too (the checker renders thousands of test scenes).
 */
#ifndef RT3_HOST_CPPDEBUGGER_HPP
#define RT3_HOST_CPPDEBUGGER_HPP

#include <cstdio>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

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
        explicit Fatal(const std::string& message) : std::runtime_error(message) {}
    };
    inline bool& verbose() { static bool flag = false; return flag; }
    inline void log(Severity severity, const std::string& message) {
        if (severity == Severity::fatal) { throw Fatal(message); }
#ifndef RT3_DEBUGGER_SILENT
        if (severity == Severity::warning || severity == Severity::nonfatal || verbose()) { std::fprintf(stderr, "[rt3] %s\n", message.c_str()); }
#endif
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
#define DLOG(SEVERITY, MESSAGE) ::CppDebugger::log((SEVERITY), (MESSAGE))

#endif
