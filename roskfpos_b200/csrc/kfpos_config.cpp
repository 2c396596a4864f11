// kfpos_config.cpp -- XML configuration semantics of the reference
// (KalmanFilter::loadConfigurationFiles, KF.cpp:749-893; attribute docs in
// config/config_pos.xml:5-26).  The reference receives each XML file's CONTENT as
// a string parameter and reads attributes of <config><uwb|px4flow|imu|mag|
// algorithm .../></config> with a default of 0 for anything absent or
// unparsable (boost::property_tree get<T>(path, default)).
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include "../../include/kfpos_b200.h"

namespace {

typedef std::map<std::string, std::string> Attrs;

double get_d(const Attrs &a, const char *k, double def = 0.0) {
    Attrs::const_iterator it = a.find(k);
    if (it == a.end()) return def;
    const char *s = it->second.c_str();
    char *end = nullptr;
    double v = strtod(s, &end);
    while (end && *end && isspace((unsigned char)*end)) ++end;
    if (end == s || (end && *end)) return def; // lexical_cast failure -> default
    return v;
}

int get_i(const Attrs &a, const char *k, int def = 0) {
    Attrs::const_iterator it = a.find(k);
    if (it == a.end()) return def;
    const char *s = it->second.c_str();
    char *end = nullptr;
    long v = strtol(s, &end, 10);
    while (end && *end && isspace((unsigned char)*end)) ++end;
    if (end == s || (end && *end)) return def;
    return (int)v;
}

void skip_ws(const char *&p) {
    while (*p && isspace((unsigned char)*p)) ++p;
}

bool is_name_char(char c) { return isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; }

// parses the attributes of the tag starting right after its name; returns
// false on a syntax error.  Leaves p after '>'.
bool parse_attrs(const char *&p, Attrs &out) {
    for (;;) {
        skip_ws(p);
        if (!*p) return false;
        if (*p == '/') {
            ++p;
            if (*p != '>') return false;
            ++p;
            return true;
        }
        if (*p == '>') { ++p; return true; }
        const char *n0 = p;
        while (is_name_char(*p)) ++p;
        if (p == n0) return false;
        std::string name(n0, p);
        skip_ws(p);
        if (*p != '=') return false;
        ++p;
        skip_ws(p);
        const char q = *p;
        if (q != '"' && q != '\'') return false;
        ++p;
        const char *v0 = p;
        while (*p && *p != q) ++p;
        if (!*p) return false;
        out[name] = std::string(v0, p);
        ++p;
    }
}

} // namespace

extern "C" void kfpos_config_default(kfpos_config *cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->ml_start[0] = 1.0; // MLLocation(... {1,1,4}), PG.cpp:531
    cfg->ml_start[1] = 1.0;
    cfg->ml_start[2] = 4.0;
}

extern "C" int kfpos_config_load_xml(kfpos_config *cfg, const char *xml) {
    if (!cfg || !xml) return KFPOS_ERR_INVALID;
    const char *p = xml;
    bool saw_config = false;
    int depth = 0;
    while (*p) {
        if (*p != '<') { ++p; continue; }
        if (!strncmp(p, "<!--", 4)) {
            const char *e = strstr(p + 4, "-->");
            if (!e) return KFPOS_ERR_PARSE;
            p = e + 3;
            continue;
        }
        if (p[1] == '?') {
            const char *e = strstr(p + 2, "?>");
            if (!e) return KFPOS_ERR_PARSE;
            p = e + 2;
            continue;
        }
        if (p[1] == '/') {
            const char *e = strchr(p, '>');
            if (!e) return KFPOS_ERR_PARSE;
            p = e + 1;
            --depth;
            continue;
        }
        ++p;
        const char *n0 = p;
        while (is_name_char(*p)) ++p;
        if (p == n0) return KFPOS_ERR_PARSE;
        const std::string name(n0, p);
        Attrs a;
        const char *before = p;
        if (!parse_attrs(p, a)) return KFPOS_ERR_PARSE;
        const bool self_closed = p - before >= 2 && p[-2] == '/';
        if (depth == 0 && name == "config") saw_config = true;
        // children of <config> only (BOOST_FOREACH over get_child("config"))
        if (depth == 1 && saw_config) {
            if (name == "px4flow") { // KF.cpp:766-779
                cfg->px4_use_fixed_sensor_height = get_i(a, "useFixedSensorHeight") == 1;
                cfg->px4_sensor_height = get_d(a, "sensorHeight");
                cfg->px4_arm_p0 = get_d(a, "armP0");
                cfg->px4_arm_p1 = get_d(a, "armP1");
                cfg->px4_sensor_init_angle = get_d(a, "sensorInitAngle");
                cfg->px4_cov_velocity = get_d(a, "covarianceVelocity");
                cfg->px4_cov_gyro_z = get_d(a, "covarianceGyroZ");
            } else if (name == "uwb") { // KF.cpp:793-800
                cfg->use_fixed_height = get_i(a, "useFixedHeight") == 1;
                cfg->fixed_height = get_d(a, "fixedHeight");
                cfg->tag_id = get_i(a, "tagId");
            } else if (name == "imu") { // KF.cpp:815-824
                cfg->imu_use_fixed_cov_acc = get_i(a, "useFixedCovarianceAcceleration") == 1;
                cfg->imu_cov_acc = get_d(a, "covarianceAcceleration");
                cfg->imu_use_fixed_cov_gyro_z = get_i(a, "useFixedCovarianceAngularVelocityZ") == 1;
                cfg->imu_cov_gyro_z = get_d(a, "covarianceAngularVelocityZ");
            } else if (name == "mag") { // KF.cpp:839-844
                cfg->mag_angle_offset = get_d(a, "angleOffset");
                cfg->mag_cov = get_d(a, "covarianceMag");
            } else if (name == "algorithm") {
                // config_pos.xml:28; the reference's parse is commented out
                // (KF.cpp:861-883) -- semantics from config_pos.xml:5-26.
                cfg->variant = get_i(a, "variant");
                cfg->num_ignored_rangings = get_i(a, "numIgnoredRangings");
                cfg->best_mode = get_i(a, "bestMode");
                cfg->min_z = get_d(a, "minZ");
                cfg->max_z = get_d(a, "maxZ");
                if (get_i(a, "useInitPosition") == 1) {
                    cfg->ml_start[0] = get_d(a, "initX");
                    cfg->ml_start[1] = get_d(a, "initY");
                    cfg->ml_start[2] = get_d(a, "initZ");
                }
            }
        }
        if (!self_closed) ++depth;
    }
    if (!saw_config) return KFPOS_ERR_PARSE; // get_child("config") would throw
    return KFPOS_OK;
}
