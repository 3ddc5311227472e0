// TEST INFRASTRUCTURE ONLY: name-only stand-ins for the UHD / HDF5 / Boost.Asio / Boost.PropertyTree declarations that the
// reference's headers/USRP_hardware_manager.hpp, USRP_file_writer.hpp, USRP_server_network.hpp and USRP_JSON_interpreter.hpp
// mention, on top of oracle/ref_stubs.  They exist so that the reference's UNCHANGED cpp/USRP_server_link_threads.cpp can be
// COMPILED (never linked or run) against include/gsdr_compat.hpp in tests/test_reference_sources_compile.py.  No behaviour.
#pragma once
#include "../../../oracle/ref_stubs/gsdr_thirdparty_stub.hpp"
#include <map>

namespace uhd {
struct time_spec_t {
    time_spec_t(double = 0.0) {}
    double get_real_secs() const { return 0.0; }
};
struct device_addr_t : std::map<std::string, std::string> {
    device_addr_t(const std::string& = "") {}
    std::string to_string() const { return ""; }
};
typedef std::vector<device_addr_t> device_addrs_t;
struct stream_args_t {
    stream_args_t(const std::string& = "", const std::string& = "") {}
    std::vector<size_t> channels;
};
struct tx_metadata_t { bool start_of_burst, end_of_burst, has_time_spec; time_spec_t time_spec; };
struct stream_cmd_t {
    enum stream_mode_t { STREAM_MODE_START_CONTINUOUS, STREAM_MODE_STOP_CONTINUOUS, STREAM_MODE_NUM_SAMPS_AND_DONE };
    stream_cmd_t(stream_mode_t) {}
    size_t num_samps; bool stream_now; time_spec_t time_spec;
};
struct rx_streamer {
    typedef std::shared_ptr<rx_streamer> sptr;
    size_t get_max_num_samps() { return 0; }
    template <class... A> size_t recv(A&&...) { return 0; }
    void issue_stream_cmd(const stream_cmd_t&) {}
};
struct tx_streamer {
    typedef std::shared_ptr<tx_streamer> sptr;
    size_t get_max_num_samps() { return 0; }
    template <class... A> size_t send(A&&...) { return 0; }
    template <class... A> bool recv_async_msg(A&&...) { return false; }
};
namespace usrp {
struct multi_usrp {
    typedef std::shared_ptr<multi_usrp> sptr;
    template <class... A> void set_time_unknown_pps(A&&...) {}
    template <class... A> void set_time_now(A&&...) {}
    time_spec_t get_time_now() { return time_spec_t(); }
};
}  // namespace usrp
}  // namespace uhd

typedef unsigned long long hsize_t;
typedef long long hid_t;
namespace H5 {
struct H5File {}; struct Group {}; struct DataSpace {}; struct DataSet {}; struct DataType {}; struct CompType {}; struct Attribute {};
struct PredType { };
struct StrType {};
}  // namespace H5

namespace boost {
namespace asio {
struct io_service {};
struct socket_base { struct reuse_address { reuse_address(bool = true) {} }; };
namespace ip {
struct address { static address from_string(const std::string&) { return address(); } };
struct tcp {
    struct endpoint { template <class... A> endpoint(A&&...) {} };
    struct socket { template <class... A> socket(A&&...) {} };
    struct acceptor { template <class... A> acceptor(A&&...) {} };
    static tcp v4() { return tcp(); }
};
}  // namespace ip
}  // namespace asio
namespace property_tree {
struct ptree {
    typedef std::string key_type;
    typedef std::pair<const std::string, ptree> value_type;
    template <class T> T get_value() const { return T(); }
    template <class T> T get(const std::string&) const { return T(); }
    ptree get_child(const key_type&) const { return ptree(); }
    const value_type* begin() const { return nullptr; }
    const value_type* end() const { return nullptr; }
};
}  // namespace property_tree
}  // namespace boost
