/* CppDebugger.hpp — the error channel of the host backend.
 *
 * The reference reports every failure through DLOG(fatal, msg), which throws
 * CppDebugger::Fatal; main() catches it and exits with -1 (reference
 * src/Main.cpp:305-308, e.g. SequentialRenderer.cpp:241,249). The backend keeps
 * that behaviour. Standalone, this header provides those names; inside the
 * reference tree the real Lut99/CppDebugger header is found first.
 */
#ifndef RT3_HOST_CPPDEBUGGER_HPP
#define RT3_HOST_CPPDEBUGGER_HPP

#include <cstdio>
#include <stdexcept>
#include <string>

namespace CppDebugger {
    enum class Severity { auxillary, info, warning, nonfatal, fatal };
    namespace SeverityValues {
        static constexpr Severity auxillary = Severity::auxillary;
        static constexpr Severity info = Severity::info;
        static constexpr Severity warning = Severity::warning;
        static constexpr Severity nonfatal = Severity::nonfatal;
        static constexpr Severity fatal = Severity::fatal;
    }
    struct Fatal : public std::runtime_error {
        explicit Fatal(const std::string& message) : std::runtime_error(message) {}
    };
    inline bool& verbose() { static bool flag = false; return flag; }
    inline void log(Severity severity, const std::string& message) {
        if (severity == Severity::fatal) { throw Fatal(message); }
        if (severity == Severity::warning || severity == Severity::nonfatal || verbose()) { std::fprintf(stderr, "[rt3] %s\n", message.c_str()); }
    }
}

#define DSTART(NAME)
#define DENTER(NAME)
#define DLEAVE
#define DRETURN return
#define DINDENT
#define DDEDENT
#define DLOG(SEVERITY, MESSAGE) ::CppDebugger::log((SEVERITY), (MESSAGE))

#endif
