# oracle/stubs/SeqToolBox.pm -- TEST INFRASTRUCTURE ONLY.
# lib/SeqToolBox.pm is missing from the reference tree although
# lib/SeqToolBox/Taxonomy.pm:31,56 needs it; this supplies the two methods used.
package SeqToolBox;
sub new       { bless {}, shift }
sub get_dbdir { $ENV{SEQTOOLBOXDB} || "." }
1;
