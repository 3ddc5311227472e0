// TEST INFRASTRUCTURE ONLY (oracle/): minimal std-backed stand-ins for the Boost and UHD
// headers that the reference's host classes include but this image does not ship.
//
// Purpose: let the reference's OWN, UNMODIFIED sources
//   /root/reference/cpp/{kernels.cu,fir.cu,USRP_demodulator.cpp,USRP_buffer_generator.cpp,
//                         USRP_server_memory_management.cpp,USRP_server_console_print.cpp}
// compile (for sm_100a) into oracle/_ref/libgsdr_ref.so, so that the real
// RX_buffer_demodulator::process / TX_buffer_generator::get run on a B200 as the parity
// reference and the "--impl reference" bench arm.  Nothing here is DSP; it only provides
// the names of third-party types those headers mention (queues, threads, log macro).
// Every stub header under oracle/ref_stubs/{boost,uhd}/ just includes this file.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <cmath>
#include <cstring>
#include <stdexcept>

namespace boost {

// ---- boost::chrono / boost::this_thread ------------------------------------------------
namespace chrono {
using std::chrono::microseconds;
using std::chrono::milliseconds;
using std::chrono::seconds;
using std::chrono::high_resolution_clock;
}  // namespace chrono

struct thread_interrupted {};

namespace detail_stub {
inline std::atomic<bool>*& current_flag() {
    static thread_local std::atomic<bool>* flag = nullptr;
    return flag;
}
}  // namespace detail_stub

namespace this_thread {
template <class Rep, class Period>
inline void sleep_for(const std::chrono::duration<Rep, Period>& d) { std::this_thread::sleep_for(d); }
inline void interruption_point() {
    std::atomic<bool>* f = detail_stub::current_flag();
    if (f && f->load()) throw thread_interrupted();
}
}  // namespace this_thread

// ---- boost::thread (std::thread + cooperative interrupt flag) ---------------------------
class thread {
  public:
    thread() = default;
    template <class F>
    explicit thread(F f) : flag_(std::make_shared<std::atomic<bool>>(false)) {
        std::shared_ptr<std::atomic<bool>> flag = flag_;
        impl_ = std::thread([flag, f]() mutable {
            detail_stub::current_flag() = flag.get();
            try { f(); } catch (thread_interrupted&) {}
        });
    }
    ~thread() { if (impl_.joinable()) impl_.detach(); }
    void interrupt() { if (flag_) flag_->store(true); }
    void join() { if (impl_.joinable()) impl_.join(); }
    bool joinable() const { return impl_.joinable(); }
    std::thread::native_handle_type native_handle() { return impl_.native_handle(); }
  private:
    std::shared_ptr<std::atomic<bool>> flag_;
    std::thread impl_;
};

template <class F, class... A>
inline auto bind(F&& f, A&&... a) -> decltype(std::bind(std::forward<F>(f), std::forward<A>(a)...)) {
    return std::bind(std::forward<F>(f), std::forward<A>(a)...);
}

using mutex = std::mutex;
using condition_variable = std::condition_variable;
template <class M> using unique_lock = std::unique_lock<M>;
template <class T> using shared_ptr = std::shared_ptr<T>;

// ---- boost::lockfree::queue (bounded, mutex-backed; same push/pop/empty contract) -------
namespace lockfree {
template <bool B> struct fixed_sized {};
template <class T, class... Opts>
class queue {
  public:
    explicit queue(size_t cap = 128) : cap_(cap) {}
    bool push(const T& v) {
        std::lock_guard<std::mutex> g(m_);
        if (q_.size() >= cap_) return false;
        q_.push_back(v);
        return true;
    }
    bool pop(T& v) {
        std::lock_guard<std::mutex> g(m_);
        if (q_.empty()) return false;
        v = q_.front();
        q_.pop_front();
        return true;
    }
    bool empty() {
        std::lock_guard<std::mutex> g(m_);
        return q_.empty();
    }
  private:
    size_t cap_;
    std::mutex m_;
    std::deque<T> q_;
};
template <class T, class... Opts> class spsc_queue : public queue<T, Opts...> {
  public:
    using queue<T, Opts...>::queue;
};
}  // namespace lockfree

// ---- boost::log names that the settings/diagnostic headers mention ----------------------
namespace log {
namespace trivial { enum severity_level { trace, debug, info, warning, error, fatal }; }
namespace sources {}
namespace keywords {}
namespace sinks {
struct text_file_backend {};
template <class B> struct synchronous_sink {};
}  // namespace sinks
}  // namespace log

}  // namespace boost

#ifndef BOOST_LOG_TRIVIAL
struct gsdr_stub_null_stream {
    template <class T> gsdr_stub_null_stream& operator<<(const T&) { return *this; }
};
#define BOOST_LOG_TRIVIAL(lvl) gsdr_stub_null_stream()
#endif

// ---- uhd names mentioned by USRP_server_diagnostic.hpp ----------------------------------
namespace uhd {
struct rx_metadata_t { enum error_code_t { ERROR_CODE_NONE = 0 }; error_code_t error_code; };
struct async_metadata_t { int event_code; };
}  // namespace uhd
