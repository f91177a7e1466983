// shim for XmlRpc (only referenced by the never-instantiated XmlRpcReader of src/utils.h:475-523)
#pragma once
#include <string>
namespace XmlRpc
{
struct XmlRpcException
{
};
struct XmlRpcValue
{
    enum Type { TypeInt, TypeDouble };
    XmlRpcValue& operator[](int) { return *this; }
    XmlRpcValue& operator[](const char*) { return *this; }
    operator bool() const { return false; }
    operator int() const { return 0; }
    operator double() const { return 0; }
    operator std::string() const { return ""; }
    Type getType() const { return TypeDouble; }
    bool hasMember(const char*) const { return false; }
};
}
